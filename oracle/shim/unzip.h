/* Build shim for the oracle/_ref recipe only (test infrastructure, never shipped).
 * minizip's compat API is used by exactly one function of the reference's tools library
 * (unzip(), tools.cpp:378-461), which is outside the hot path; every call fails cleanly. */
#ifndef ORACLE_SHIM_UNZIP_H
#define ORACLE_SHIM_UNZIP_H
typedef void* unzFile;
typedef struct { unsigned long number_entry; unsigned long size_comment; } unz_global_info;
typedef struct { unsigned long uncompressed_size; unsigned long compressed_size; unsigned long size_filename; } unz_file_info;
#define UNZ_OK 0
#define UNZ_END_OF_LIST_OF_FILE (-100)
static inline unzFile unzOpen(const char* p) { (void)p; return (unzFile)0; }
static inline int unzGetGlobalInfo(unzFile f, unz_global_info* i) { (void)f; (void)i; return -1; }
static inline int unzGoToFirstFile(unzFile f) { (void)f; return -1; }
static inline int unzGetCurrentFileInfo(unzFile f, unz_file_info* i, char* n, unsigned long ns, void* e, unsigned long es, char* c, unsigned long cs) { (void)f; (void)i; (void)n; (void)ns; (void)e; (void)es; (void)c; (void)cs; return -1; }
static inline int unzOpenCurrentFile(unzFile f) { (void)f; return -1; }
static inline int unzReadCurrentFile(unzFile f, void* b, unsigned n) { (void)f; (void)b; (void)n; return -1; }
static inline int unzCloseCurrentFile(unzFile f) { (void)f; return -1; }
static inline int unzGoToNextFile(unzFile f) { (void)f; return UNZ_END_OF_LIST_OF_FILE; }
static inline int unzClose(unzFile f) { (void)f; return 0; }
#endif
