// qpde_kernels.cuh — device-side views shared between the C-ABI layer
// (qpde_api.cu) and the kernel families.
#ifndef QPDE_KERNELS_CUH
#define QPDE_KERNELS_CUH

#include <cuda_runtime.h>

#include <string>

#include "qpde_common.cuh"

namespace qpde {

// Device copy of qpde_bs1d_batch (all pointers are device pointers).
struct DeviceBatch {
	int C, N;
	int c0;                // RESIDENT family: first contract of this launch (C of them are swept; arrays are indexed by contract)
	const double *axis; long long axis_stride; int n_axis, refine;
	const double *r, *sigma, *q;
	int vol_model;
	const double *sigma_nodes; long long sigma_nodes_stride;
	const double *T, *dt;
	int scheme, max_steps;
	int forward;           // 1: forward-time classes (time runs from 0 to T)
	const double *variable_target; double variable_scale;
	int n_exercise, exercise_kind;
	int dividend_kind, dividend_first;
	const double *dividend;
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
	const double *jump_lambda, *jump_kappa, *jump_x0, *jump_dx, *jump_weights;
	long long jump_weights_stride; int n_fft;
	// general events (QPDE_EV_* programs); event_program == nullptr: the tagged exercise / dividend transforms
	const int *event_program; const double *event_consts, *event_params;
	long long event_params_stride; int n_event_params;
	const double *event_times;      // [n_exercise] ascending, or nullptr: uniform T e / E
	const double *source_table; int n_source;   // b(t) = source_table[floor(t)] S, or nullptr
};

// Device output buffers (contract-major); any pointer may be null.
struct DeviceResult {
	double *V; long long V_stride;
	double *S; long long S_stride;
	double *value;
	int *n_steps;
	int *inner_total;
	int *computed;         // inner solves actually executed (inner_total also counts the confirming ones)
	unsigned char *inner; long long inner_stride;
	unsigned char *active; long long active_stride;
	int *status;
};

// Workspace of the INTERLEAVED family: arrays of N x ld doubles, ld = C rounded up to a warp of
// contracts, laid out [ld / 32][N][32]: contract c, node i lives at ((c / 32) N + i) 32 + c % 32.
constexpr int QPDE_WARP_ROW = 32;
QPDE_HD size_t interleaved_base(int c, int N) { return (size_t)(c >> 5) * (size_t)N * QPDE_WARP_ROW + (size_t)(c & 31); }
struct InterleavedWork {
	double *S, *lo, *di, *up;
	double *x, *xprev, *lb, *cp, *dp;
	double *obst, *evobst; // null when unused
	long long ld;          // contracts incl. padding (an array is N x ld doubles)
};

// The TMA-staged sweep (qpde_interleaved_tma.cu) reads every array through tiled tensor maps over the
// workspace block seen as [array][rows][32]: one map per box depth (1, 2 or 3 adjacent arrays).
struct InterleavedTma {
	alignas(64) unsigned char map[3][128];   // CUtensorMap x 3 (opaque here; cuda.h stays out of this header)
	bool enabled;
	int stages;                           // ring depth (tiles in flight per warp)
};
constexpr int QPDE_TMA_TILE_ROWS = 4;     // node rows per TMA tile (box = 32 contracts x 4 rows = 1 KB)

// true if the TMA-staged sweep covers this batch (N >= 32, schemes up to BDF2, no policy iteration; dividend
// events and the variable stepper run inside it with ordinary loads)
bool interleaved_tma_supported(const DeviceBatch &b);
// encodes the tensor maps over `block` (`arrays` arrays of `rows` x 32 doubles); false if the driver
// entry point is missing
bool interleaved_tma_prepare(InterleavedTma *t, double *block, long long rows, int arrays, std::string *error);
void launch_interleaved_tma(const DeviceBatch &b, const InterleavedWork &w, const InterleavedTma &t, double *block,
		const DeviceResult &out, cudaStream_t stream, long long *launches);
void launch_interleaved_setup(const DeviceBatch &b, const InterleavedWork &w, cudaStream_t stream, long long *launches);

void launch_interleaved(const DeviceBatch &b, const InterleavedWork &w,
		const DeviceResult &out, cudaStream_t stream, long long *launches);

// Workspace of the jump-diffusion kernel (qpde_jump.cu), per contract.
struct JumpWork {
	int *g_idx; double *g_w;      // [C][n_fft]: V(exp(x0 + k dx)) = w V[idx] + (1 - w) V[idx + 1]
	int *n_idx; double *n_w;      // [C][N]:     h(log S_i)       = w h[idx] + (1 - w) h[idx + 1]
};
bool jump_supported(const DeviceBatch &b);
int launch_jump(const DeviceBatch &b, const JumpWork &w, const DeviceResult &out,
		cudaStream_t stream, long long *launches);

// The communicator of a handle (qpde_comm.cu): NCCL, loaded at run time.  Only the sharded GMWB solve uses it.
struct Comm;
int comm_unique_id(unsigned char *id, std::string *error);
int comm_create(const unsigned char *id, int rank, int world, Comm **out, std::string *error);
void comm_destroy(Comm *c);
int comm_rank(const Comm *c);
int comm_world(const Comm *c);
bool comm_group_start(Comm *c);
bool comm_group_end(Comm *c);
bool comm_broadcast(Comm *c, void *buf, size_t bytes, int root, cudaStream_t stream);
bool comm_allreduce_max(Comm *c, int *buf, int count, cudaStream_t stream);

// GMWB 2-D sweep (qpde_gmwb.cu): host-driven loop over the kernels.  comm != nullptr: the control search
// is sharded by the A-axis over the communicator's ranks (a collective call).
int gmwb_run(const qpde_gmwb2d *p, const double *W, int nw, const double *A, int na, const double *z, int nz,
		double *V_host, qpde_gmwb2d_stats *stats, cudaStream_t stream, long long *launches, std::string *error,
		Comm *comm = nullptr);

struct GmwbShard;
int gmwb_shard_create(const qpde_gmwb2d *p, const double *W, int nw, const double *A, int na, const double *z, int nz,
		int a_begin, int a_end, cudaStream_t stream, long long *launches, GmwbShard **out, std::string *error);
void gmwb_shard_destroy(GmwbShard *sh);
void gmwb_shard_exit(GmwbShard *sh, double *x);
void gmwb_shard_policy(GmwbShard *sh, const double *x, double h0);
cudaError_t gmwb_shard_sweep(GmwbShard *sh, double *x, const double *v0, int *status_dev);
void gmwb_shard_relerr(GmwbShard *sh, const double *x, const double *xprev, int *notdone_dev);

// 1-D impulse-control HJBQVI (qpde_hjbqvi1d.cu): one CTA per problem.  Device pointers.
struct Hjbqvi1dBatch {
	int n_problems, model, N, nq, nz;
	int refine, refine_q, refine_z;
	const double *x_axis, *q_axis, *z_axis; long long x_axis_stride, q_axis_stride, z_axis_stride;   // coarse axes per problem (stride 0: shared)
	const double *params; long long params_stride;
	double T; int timesteps, max_inner, max_sweeps;
	double scaling_factor, tolerance;
};
struct Hjbqvi1dResult {
	double *V, *X, *Q, *Z;          // [n_problems][N]; any may be null
	unsigned char *mask;
	double *mean_inner, *mean_sweeps; int *n_steps, *status;
};
size_t hjbqvi1d_shared_bytes(int N, int nq, int nz);
int launch_hjbqvi1d(const Hjbqvi1dBatch &b, const Hjbqvi1dResult &out, cudaStream_t stream, long long *launches);

// 2-D impulse-control HJBQVI with impulses both ways (qpde_hjbqvi2d.cu): host-driven loop over the kernels.
int hjbqvi2d_run(const qpde_hjbqvi2d *p, const double *W, int nw, const double *B, int nb, const double *q, int nq,
		const double *z, int nz, double *V_host, double *Q_host, double *Z_host, unsigned char *mask_host,
		qpde_hjbqvi2d_stats *stats, cudaStream_t stream, long long *launches, std::string *error);

// FP64 pipe rate of this device, measured (qpde_peak.cu)
int measure_fp64_peak(int sm_count, cudaStream_t stream, double *tflops, long long *launches);

// RESIDENT family at two contracts per SM (qpde_lean.cu): American sweeps on 513 < N <= 2049 nodes.
bool lean_supported(const DeviceBatch &b);
int launch_lean(const DeviceBatch &b, const DeviceResult &out, cudaStream_t stream, long long *launches);

// RESIDENT family (qpde_resident.cu): one CTA per contract.
bool resident_supported(const DeviceBatch &b);
int launch_resident(const DeviceBatch &b, const DeviceResult &out,
		cudaStream_t stream, int sm_count, long long *launches);

} // namespace qpde

#endif
