// qpde_kernels.cuh — device-side views shared between the C-ABI layer
// (qpde_api.cu) and the kernel families.
#ifndef QPDE_KERNELS_CUH
#define QPDE_KERNELS_CUH

#include <cuda_runtime.h>

#include "qpde_common.cuh"

namespace qpde {

// Device copy of qpde_bs1d_batch (all pointers are device pointers).
struct DeviceBatch {
	int C, N;
	const double *axis; long long axis_stride; int n_axis, refine;
	const double *r, *sigma, *q;
	int vol_model;
	const double *sigma_nodes; long long sigma_nodes_stride;
	const double *T, *dt;
	int scheme, max_steps;
	int n_exercise, exercise_kind;
	const double *strike;
	int payoff_kind, obstacle_kind;
	const double *payoff; long long payoff_stride;
	const double *obstacle; long long obstacle_stride;
	int american, max_inner;
	double tolerance, scale, penalty_tolerance;
	double pen;   // 1 / penalty_tolerance (PenaltyMethod.hpp:85), formed on the host
	const double *spot;
	int n_controls, policy_max;
	const double *controls; long long controls_stride;
};

// Device output buffers (contract-major); any pointer may be null.
struct DeviceResult {
	double *V; long long V_stride;
	double *S; long long S_stride;
	double *value;
	int *n_steps;
	int *inner_total;
	unsigned char *inner; long long inner_stride;
	unsigned char *active; long long active_stride;
	int *status;
};

// Workspace of the INTERLEAVED family: [N][C] arrays.
struct InterleavedWork {
	double *S, *lo, *di, *up;
	double *x, *xprev, *lb, *cp, *dp;
	double *obst, *evobst; // null when unused
};

void launch_interleaved(const DeviceBatch &b, const InterleavedWork &w,
		const DeviceResult &out, cudaStream_t stream, long long *launches);

// RESIDENT family (qpde_resident.cu): one CTA per contract.
bool resident_supported(const DeviceBatch &b);
int launch_resident(const DeviceBatch &b, const DeviceResult &out,
		cudaStream_t stream, int sm_count, long long *launches);

} // namespace qpde

#endif
