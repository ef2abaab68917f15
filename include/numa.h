// numa.h -- stand-in for libnuma's header.
//
// The reference's driver includes <numa.h> (main.cpp:10) and its link line ends in -lnuma
// (Makefile:11) although main.cpp itself calls nothing from libnuma.  This image has neither the
// header nor the library, so the build ships this header (found through -I./include) and a
// stub libnuma.a (bin/libnuma.a, found through LIBRARY_PATH).  On a machine with the real
// libnuma, delete this file and the real one is used; the GPU library does not need it.
#ifndef THSP_NUMA_SHIM_H
#define THSP_NUMA_SHIM_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
int   numa_available(void);
int   numa_num_configured_nodes(void);
void* numa_alloc_onnode(size_t size, int node);
void  numa_free(void* start, size_t size);
int   numa_run_on_node(int node);
#ifdef __cplusplus
}
#endif
#endif
