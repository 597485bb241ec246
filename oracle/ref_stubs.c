/*
 * ref_stubs.c — empty stand-ins for the shared libraries the reference's librocksdb.so.5.4 names in
 * DT_NEEDED but this image lacks (snappy, gflags, zstd, numa, jemalloc, hdfs, JVM bits).
 * TEST INFRASTRUCTURE ONLY.  Compression is forced to kNoCompression by ref_driver.c, HDFS/JVM paths
 * are never taken; every stub other than the two called from static initialisers aborts if reached.
 */
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>

#define DIE(name)                                                         \
  void name(void) {                                                       \
    fprintf(stderr, "oracle/_ref stub reached: %s\n", #name);             \
    abort();                                                              \
  }

unsigned ZSTD_versionNumber(void) { return 10102; } /* called from a static initialiser */
void malloc_stats_print(void (*w)(void*, const char*), void* o, const char* opts) {
  (void)w; (void)o; (void)opts;
}
DIE(ZSTD_compressBound) DIE(ZSTD_compress_usingDict) DIE(ZSTD_createCCtx) DIE(ZSTD_createDCtx)
DIE(ZSTD_decompress_usingDict) DIE(ZSTD_freeCCtx) DIE(ZSTD_freeDCtx)
DIE(_ZN6snappy11RawCompressEPKcmPcPm) DIE(_ZN6snappy13RawUncompressEPKcmPc)
DIE(_ZN6snappy19MaxCompressedLengthEm) DIE(_ZN6snappy21GetUncompressedLengthEPKcmPm)
DIE(hdfsCloseFile) DIE(hdfsConnectNewInstance) DIE(hdfsCreateDirectory) DIE(hdfsDelete)
DIE(hdfsDisconnect) DIE(hdfsExists) DIE(hdfsFlush) DIE(hdfsFreeFileInfo) DIE(hdfsGetPathInfo)
DIE(hdfsHSync) DIE(hdfsListDirectory) DIE(hdfsOpenFile) DIE(hdfsPread) DIE(hdfsRead) DIE(hdfsRename)
DIE(hdfsSeek) DIE(hdfsTell) DIE(hdfsWrite)
