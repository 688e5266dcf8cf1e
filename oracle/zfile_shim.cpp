// oracle/zfile_shim.cpp -- TEST INFRASTRUCTURE ONLY.
// extern "C" doors onto the reference's zstd movie file (ZFile.h: C++ linkage, a FileReaderPtr argument),
// compiled by build_ref.sh together with ZFile.cpp where it lies under /root/reference and linked against
// the compiled reference tools library (FileAttributes, zstd wrappers).  Nothing of the reference is copied.
#include "rir_config.h"
#include "ReadFileChunk.h"
#include "ZFile.h"

#define DOOR extern "C" __attribute__((visibility("default")))

DOOR void* ref_z_open_file_write(const char* filename, int w, int h, int rate, int method, int clevel)
{
    return z_open_file_write(filename, w, h, rate, method, clevel);
}
DOOR void* ref_z_open_file_read(const char* filename)
{
    return z_open_file_read(rir::createFileReader(rir::createFileAccess(filename)));
}
DOOR unsigned long long ref_z_close_file(void* f) { return z_close_file(f); }
DOOR int ref_z_image_count(void* f) { return z_image_count(f); }
DOOR int ref_z_image_size(void* f, int* w, int* h) { return z_image_size(f, w, h); }
DOOR int ref_z_write_image(void* f, const unsigned short* img, long long ts) { return z_write_image(f, img, ts); }
DOOR int ref_z_read_image(void* f, int pos, unsigned short* img, long long* ts)
{
    int64_t t = 0;
    int r = z_read_image(f, pos, img, &t);
    if (ts) *ts = t;
    return r;
}
DOOR int ref_z_get_timestamps(void* f, long long* out)
{
    int n = z_image_count(f);
    int64_t* t = z_get_timestamps(f);
    if (!t) return -1;
    for (int i = 0; i < n; ++i) out[i] = t[i];
    return n;
}
