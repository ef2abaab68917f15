/* libnuma.a stand-in for images without libnuma (see include/numa.h).  One node, malloc-backed. */
#include <stdlib.h>
int numa_available(void) { return 0; }
int numa_num_configured_nodes(void) { return 1; }
void* numa_alloc_onnode(size_t size, int node) { (void)node; return malloc(size ? size : 1); }
void numa_free(void* start, size_t size) { (void)size; free(start); }
int numa_run_on_node(int node) { (void)node; return 0; }
