/* Stand-in for <numa.h> when compiling the reference for the oracle: libnuma is not
 * installed in this image.  The reference calls exactly these four functions
 * (src/mat_vec.cpp:150,188-192,220-224,490 and the four sibling blocks). */
#ifndef ORACLE_SHIM_NUMA_H
#define ORACLE_SHIM_NUMA_H
#include <stdlib.h>
static inline int numa_num_configured_nodes(void) { return 1; }
static inline void* numa_alloc_onnode(size_t bytes, int node) { (void)node; return malloc(bytes ? bytes : 1); }
static inline void numa_free(void* p, size_t bytes) { (void)bytes; free(p); }
static inline int numa_run_on_node(int node) { (void)node; return 0; }
#endif
