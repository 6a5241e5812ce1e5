// Host build of csrc/pnp_core.inl (the minimal-solver arithmetic the PnP-RANSAC kernel executes), exposed so the CPU
// test-suite can check it against numpy before anything runs on a GPU.  Test infrastructure only.
#define SB_HOST_MODEL
#include "../a-simple-stereo-slam-system-with-deep-loop-closing_b200/csrc/pnp_core.inl"

extern "C" {
int hm_quartic(double a, double b, double c, double d, double *x) { return pnp_quartic(a, b, c, d, x); }
int hm_p3p(const double *P, const double *j, double *Rt) { return pnp_p3p(P, j, Rt); }
}
