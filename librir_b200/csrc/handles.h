// librir_b200/csrc/handles.h -- integer handle tables for the objects the C ABI hands out.
#pragma once
#include <map>
#include <memory>
#include <mutex>

namespace rirb {

// Lowest free positive id, like the reference's set_void_ptr (tools.cpp:40-85); mutex-guarded.
template <typename T> struct Table {
    std::mutex mu;
    std::map<int, std::shared_ptr<T>> items;
    int add(const std::shared_ptr<T>& p)
    {
        std::lock_guard<std::mutex> lock(mu);
        int id = 1;
        for (auto& kv : items) {
            if (kv.first != id) break;
            ++id;
        }
        items[id] = p;
        return id;
    }
    std::shared_ptr<T> get(int id)
    {
        std::lock_guard<std::mutex> lock(mu);
        auto it = items.find(id);
        return it == items.end() ? nullptr : it->second;
    }
    void remove(int id)
    {
        std::lock_guard<std::mutex> lock(mu);
        items.erase(id);
    }
};

}  // namespace rirb
