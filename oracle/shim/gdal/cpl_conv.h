/* TEST INFRASTRUCTURE ONLY -- see gdal_priv.h in this directory. */
#include "gdal_priv.h"
