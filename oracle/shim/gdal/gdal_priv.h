/* TEST INFRASTRUCTURE ONLY -- just enough of the GDAL C++ API for the reference's
 * 3rdparty/homography/LibImages/LibImages.cpp to compile without libgdal (which is not installed here).
 * Only Image::readGDAL / Image::writeGDAL touch these symbols and the oracle harness never calls them:
 * it feeds the reference's runHomography() from memory (oracle/homography_harness.cpp). */
#ifndef S2PB_GDAL_STUB_H
#define S2PB_GDAL_STUB_H
#include <cstddef>
#include <cstdlib>
typedef void *GDALDatasetH;
typedef int CPLErr;
enum { CPLE_None = 0 };
enum GDALAccess { GA_ReadOnly = 0 };
enum GDALRWFlag { GF_Read = 0, GF_Write = 1 };
enum GDALDataType { GDT_Float32 = 6 };
#ifndef FALSE
#define FALSE 0
#endif
struct GDALRasterBand {
    CPLErr RasterIO(GDALRWFlag, int, int, int, int, void *, int, int, GDALDataType, int, int) { return 1; }
};
struct GDALDataset {
    int GetRasterXSize() { return 0; }
    int GetRasterYSize() { return 0; }
    int GetRasterCount() { return 0; }
    GDALRasterBand *GetRasterBand(int) { return NULL; }
};
struct GDALDriver {
    GDALDataset *Create(const char *, int, int, int, GDALDataType, char **) { return NULL; }
    GDALDataset *CreateCopy(const char *, GDALDataset *, int, char **, void *, void *) { return NULL; }
};
struct GDALDriverManager { GDALDriver *GetDriverByName(const char *) { return NULL; } };
static inline GDALDriverManager *GetGDALDriverManager() { static GDALDriverManager m; return &m; }
static inline void GDALAllRegister() {}
static inline GDALDatasetH GDALOpen(const char *, GDALAccess) { return NULL; }
static inline void GDALClose(GDALDatasetH) {}
static inline void GDALDestroyDriverManager() {}
static inline void *CPLMalloc(size_t n) { return malloc(n); }
static inline void CPLFree(void *p) { free(p); }
#endif
