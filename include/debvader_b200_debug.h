/*
 * debvader_b200 — debug / ablation entry points.  They exist ONLY in the ablation build of the library
 * (libdebvader_b200_ablate.so, compiled with -DDBV_ABLATE by `python -m debvader_b200._build --ablate`); the product
 * library (libdebvader_b200.so) exports none of them and reads no DBV_* environment switch.
 */
#ifndef DEBVADER_B200_DEBUG_H
#define DEBVADER_B200_DEBUG_H

#include "debvader_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* tcgen05 / TMA self-test kernels (descriptor conventions this library relies on).  `which`
 * selects the probe; out_dev receives the kernel result, see csrc/tc_probe.cu. */
int dbv_probe(int which, const void* a_dev, const void* b_dev, float* out_dev, int M, int N, int K, void* stream);

/* clock64 instrumentation of the resident-halo kernel (csrc/tc_halo.cu): per-role cycle counters summed over the CTAs
 * of the launches made since the last reset.  out[0..DBV_HALO_NCOUNTERS) as documented in tools/halo_clocks.py. */
#define DBV_HALO_NCOUNTERS 16
#define DBV_HALO_NLAYERS 24
/* out_host: [DBV_HALO_NLAYERS][DBV_HALO_NCOUNTERS], indexed by the layer's position in the network (0 = enc_conv1 ... 19 = dec_head) */
int dbv_halo_counters(unsigned long long* out_host, int reset);

#ifdef __cplusplus
}
#endif
#endif /* DEBVADER_B200_DEBUG_H */
