/* Forwarder so that sources written against the reference (`#include "cwipc_util/api.h"`, e.g. the
 * cwipc_downsample / cwipc_remove_outliers / cwipc_tilefilter apps) compile unchanged against
 * libcwipc_util_cuda.  The declarations live in ../cwipc_util_cuda.h. */
#ifndef CWIPC_UTIL_CUDA_API_FORWARD_H
#define CWIPC_UTIL_CUDA_API_FORWARD_H
#include "../cwipc_util_cuda.h"
#endif
