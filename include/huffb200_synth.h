/* huffb200_synth.h — bench/test support exported by libhuffb200.so.  NOT part of the drop-in
 * boundary (the reference's util/TestDataGenerator.java is out of scope); it exists so that
 * bench.py and the GPU tests can fill multi-GiB device buffers with a reproducible stream. */
#ifndef HUFFB200_SYNTH_H
#define HUFFB200_SYNTH_H
#include "huffb200.h"
#ifdef __cplusplus
extern "C" {
#endif
/* d_out[i] = qtable65536[ mix64(seed, stream_offset + i) >> 48 ] for i in [0, n), where
 *   mix64(s, i): z = s + (i+1)*0x9E3779B97F4A7C15; z = (z ^ z>>30)*0xBF58476D1CE4E5B9;
 *                z = (z ^ z>>27)*0x94D049BB133111EB; return z ^ z>>31            (splitmix64)
 * d_out must be device memory; qtable65536 may be host or device memory. */
int hz_synth_fill(hz_ctx* ctx, uint8_t* d_out, uint64_t n, uint64_t stream_offset, uint64_t seed,
                  const uint8_t* qtable65536);
/* Re-read the developer knobs (HZ_* environment variables that select kernel variants for A/B runs and for the
 * parity tests of the alternative paths).  They are otherwise read once, in hz_create. */
int hz_dev_reload_knobs(hz_ctx* ctx);
#ifdef __cplusplus
}
#endif
#endif
