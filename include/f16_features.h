/* f16_features.h - C ABI of the per-frame feature transform (libf16b200.so), SURVEY.md 8(f) row 2.
 *
 * Replaces JSBSimFeatureExtractor.forward (jsbsim_gym/features.py:37-67), which the reference applies to
 * every frame of the stacked observation (jsbsim_gym/LMA_features.py:744-771): 15 raw features
 * [x, y, h, mach, alpha, beta, p, q, r, phi, theta, psi, gx, gy, gz] -> 17 features
 * [1/(1+d*1e-3), dz/15000, h/15000, mach, p, q, r, cos a, cos b, sin a, sin b, cos phi, cos theta,
 *  sin phi, sin theta, cos rel_bearing, sin rel_bearing], d = horizontal distance to the goal,
 * rel_bearing = atan2(dy, dx) - psi. Device pointers, caller's stream. */
#ifndef F16_FEATURES_H
#define F16_FEATURES_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
/* frames: [n_frames][15] float32 -> out: [n_frames][17] float32 (n_frames = batch * 10 for stacked obs). */
int f16_features17(int64_t n_frames, const float* frames, float* out, void* stream);
#ifdef __cplusplus
}
#endif
#endif
