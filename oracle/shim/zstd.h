/* Build shim for the oracle/_ref recipe only (test infrastructure, never shipped).
 * The reference's tools library calls five zstd entry points (tools.cpp:352-376,
 * FileAttributes.cpp:67-68,134); the image has libzstd.so.1 (v1.5.5, the version the
 * reference pins in extra/CMakeLists.txt:20-31) but no header, so declare them here. */
#ifndef ORACLE_SHIM_ZSTD_H
#define ORACLE_SHIM_ZSTD_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
size_t ZSTD_compressBound(size_t srcSize);
unsigned ZSTD_isError(size_t code);
unsigned long long ZSTD_getFrameContentSize(const void* src, size_t srcSize);
size_t ZSTD_compress(void* dst, size_t dstCapacity, const void* src, size_t srcSize, int compressionLevel);
size_t ZSTD_decompress(void* dst, size_t dstCapacity, const void* src, size_t compressedSize);
#define ZSTD_CONTENTSIZE_UNKNOWN (0ULL - 1)
#define ZSTD_CONTENTSIZE_ERROR (0ULL - 2)
#ifdef __cplusplus
}
#endif
#endif
