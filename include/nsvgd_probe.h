/* nsvgd_probe.h — tcgen05 bring-up probes, TEST-ONLY (libnsvgd_probe.so; tests/test_tc_probe.py).  Not part of the
 * drop-in API and not linked into libnsvgd.so. */
#ifndef NSVGD_PROBE_H_
#define NSVGD_PROBE_H_
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
int nsvgd_tc_probe_gram(void* stream, const float* A /*128 x 128*/, const float* B /*64 x 128*/,
                        float* G /*128 x 64*/, void* ws, size_t ws_bytes);
int nsvgd_tc_probe_pv(void* stream, const float* P /*128 x 64*/, const float* V /*64 x 256*/,
                      float* O /*128 x 256*/, void* ws, size_t ws_bytes);
#ifdef __cplusplus
}
#endif
#endif
