/*
 * bnpc_b200_debug.h -- debugging hooks of libbnpc_b200.so.  NOT part of the drop-in ABI
 * (include/bnpc_b200.h): process-wide state, not thread safe, used by tools/ only.
 */
#ifndef BNPC_B200_DEBUG_H
#define BNPC_B200_DEBUG_H
#ifdef __cplusplus
extern "C" {
#endif
/* buf = device array of 4096 int64 (or NULL) that CTA 0 of bnpc_ll_matrix_i8 fills with clock64
 * stamps of its pipeline phases (tools/tc_trace.py).  One pointer per process: set it, launch on
 * ONE stream, read it, clear it.                                                              */
int bnpc_debug_set_trace(void* buf);
#ifdef __cplusplus
}
#endif
#endif
