// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE ONLY.
// extern "C" doors onto two header-only templates of the reference (Filters.h) that its C
// facade does not expose directly, so that the C restatement can be checked against the
// reference's own code for them.  Compiled by build_ref.sh against the headers where they
// lie under /root/reference; nothing of the reference is copied.
//   ref_bad_pixels_list   -> rir::badPixels<unsigned short>      (Filters.h:135-193)
//   ref_translate_u16_f32 -> rir::translate<unsigned short,float> (Filters.h:249-326),
//                            the instantiation removeMotionGeneric uses (IRFileLoader.cpp:617-627)
#include "rir_config.h"
#include "Filters.h"
#include <algorithm>
#include <vector>

extern "C" __attribute__((visibility("default")))
int ref_bad_pixels_list(const unsigned short* img, int w, int h, double std_factor, int* xy, int cap)
{
    rir::Polygon p = rir::badPixels(img, (size_t)w, (size_t)h, std_factor);
    int n = (int)p.size();
    for (int i = 0; i < n && i < cap; ++i) {
        xy[2 * i] = (int)p[i].x();
        xy[2 * i + 1] = (int)p[i].y();
    }
    return n;
}

extern "C" __attribute__((visibility("default")))
void ref_remove_motion(unsigned short* img, int w, int h, double sx, double sy)
{
    std::vector<float> tmp((size_t)w * h);
    rir::translate(img, tmp.data(), 0.f, (size_t)w, (size_t)h, (float)-sx, (float)-sy, rir::TranslateNearest);
    std::copy(tmp.begin(), tmp.end(), img);
}
