// kfbo_kernels.h -- host-visible launch interface of kfbo_kernels.cu.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

namespace kfbo {

struct Case;  // kfbo_device.cuh

// d_cases[ncases]; d_gu[D][kMaxSteps] = (g0*u, g1*u); trials [trial0, trial0+ntrials).
// d_per_trial (optional) is [slots][4][per_trial_ld]; d_partials [ncases][tiles][4];
// d_tickets [ncases] zero-initialised once; d_sums [slots][4].
cudaError_t launch_rollout_philox(const Case *d_cases, int ncases, const double2 *d_gu, int max_T, uint64_t seed,
                                  uint32_t trial0, uint32_t ntrials, double *d_per_trial, size_t per_trial_ld,
                                  double *d_partials, unsigned *d_tickets, double *d_sums, cudaStream_t stream);

cudaError_t launch_rollout_supplied(const Case *d_cases, int ncases, const double2 *d_gu, int max_T, const double *d_x0noise,
                                    const double *d_w, const double *d_v, size_t B, double *d_per_trial,
                                    double *d_partials, unsigned *d_tickets, double *d_sums, cudaStream_t stream);

cudaError_t launch_dfma_chain(double *d_out, int blocks, int iters, cudaStream_t stream);

int rollout_tiles(size_t ntrials);

}  // namespace kfbo
