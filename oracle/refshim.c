/* refshim.c — checker-side glue compiled INTO oracle/_ref/libbwa_ref.so only.
 * Satisfies the one extern the reference library expects from its host program
 * (bwa_pg: /root/reference/go/src/gobwa/bwa_bridge.c:12) and exposes sizeof probes used by tests.
 * TEST INFRASTRUCTURE: never linked into libcuda_lariat.so. */
#include <stddef.h>
#include "bwamem.h"
#include "bwa.h"
char *bwa_pg = "cuda-lariat oracle";
int refshim_sizeof_alnreg(void) { return (int)sizeof(mem_alnreg_t); }
int refshim_sizeof_opt(void) { return (int)sizeof(mem_opt_t); }
int refshim_sizeof_aln(void) { return (int)sizeof(mem_aln_t); }
