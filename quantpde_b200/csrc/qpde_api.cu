// qpde_api.cu — the C ABI of libqpde_b200.so (include/qpde_b200.h): context,
// device-resident plans, host<->device staging and kernel dispatch.  No torch,
// no C++ types across the boundary.  There is no CPU fallback: without an
// sm_100 device qpde_create() fails with QPDE_ERR_NO_DEVICE.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <mutex>
#include <string>
#include <vector>

#include "qpde_kernels.cuh"

using namespace qpde;

namespace {

std::mutex g_err_mutex;
std::string g_last_error;

void set_global_error(const std::string &s) {
	std::lock_guard<std::mutex> lock(g_err_mutex);
	g_last_error = s;
}

} // namespace

// Device blocks of destroyed plans, kept for the next plan of this handle: qpde_bs1d_sweep() makes a
// plan per call, and cudaMalloc / cudaFree of its gigabyte-sized outputs cost 15 - 180 ms per call
// (cudaFree also waits for the device).  Everything on a handle runs on its one stream, so a recycled
// block is ordered behind its last use.  Only blocks up to POOL_MAX_BLOCK are kept, up to
// POOL_MAX_TOTAL in all (the multi-gigabyte INTERLEAVED workspaces go back to the driver).
struct DevicePool {
	static constexpr size_t POOL_MAX_BLOCK = (size_t)2 << 30, POOL_MAX_TOTAL = (size_t)6 << 30;
	std::vector<std::pair<size_t, void *>> blocks;   // (bytes, pointer)
	size_t bytes = 0;
	void *take(size_t want) {
		size_t best = blocks.size();
		for(size_t k = 0; k < blocks.size(); ++k) {
			const size_t have = blocks[k].first;
			if(have >= want && have <= want + want / 8 + 4096 && (best == blocks.size() || have < blocks[best].first)) best = k;
		}
		if(best == blocks.size()) return nullptr;
		void *ptr = blocks[best].second;
		bytes -= blocks[best].first;
		blocks.erase(blocks.begin() + (long)best);
		return ptr;
	}
	void give(size_t size, void *ptr) {
		if(size > POOL_MAX_BLOCK || bytes + size > POOL_MAX_TOTAL) { cudaFree(ptr); return; }
		blocks.emplace_back(size, ptr);
		bytes += size;
	}
	void release() {
		for(auto &blk : blocks) cudaFree(blk.second);
		blocks.clear(); bytes = 0;
	}
};

struct qpde_handle {
	int device;
	int sm_count;
	cudaStream_t stream;
	cudaStream_t copy_stream;   // device-to-host copies of finished contract ranges, behind the sweep of the next range
	long long launches;
	std::string error;
	DevicePool pool;
	Comm *comm = nullptr;       // qpde_comm_init: the ranks of a sharded GMWB solve
};

struct qpde_bs1d_plan {
	qpde_handle *h;
	DeviceBatch b;
	DeviceResult out;
	InterleavedWork w;
	InterleavedTma tma;
	double *wblock;   // the one allocation behind the INTERLEAVED arrays when the TMA-staged sweep is used
	JumpWork jw;
	bool jump;
	int kernel;
	int outputs;
	std::vector<std::pair<size_t, void *>> allocations;   // (bytes, pointer)
};

namespace {

int fail(qpde_handle *h, int code, const char *fmt, ...) {
	char buf[512];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof(buf), fmt, ap);
	va_end(ap);
	if(h) h->error = buf;
	set_global_error(buf);
	return code;
}

#define QPDE_CUDA(h, call) \
	do { \
		cudaError_t qpde_err_ = (call); \
		if(qpde_err_ != cudaSuccess) { \
			return fail((h), QPDE_ERR_CUDA, "%s failed: %s (%s:%d)", #call, \
				cudaGetErrorString(qpde_err_), __FILE__, __LINE__); \
		} \
	} while(0)

template <typename T>
int dev_alloc(qpde_bs1d_plan *p, T **out, size_t count) {
	size_t bytes = count * sizeof(T) > 0 ? count * sizeof(T) : 8;
	bytes = (bytes + 255) / 256 * 256;
	void *ptr = p->h->pool.take(bytes);
	if(!ptr) {
		cudaError_t err = cudaMalloc(&ptr, bytes);
		if(err != cudaSuccess && !p->h->pool.blocks.empty()) {
			// out of memory with blocks parked in the pool: give them back and try again
			(void)cudaGetLastError();
			p->h->pool.release();
			err = cudaMalloc(&ptr, bytes);
		}
		QPDE_CUDA(p->h, err);
	}
	p->allocations.emplace_back(bytes, ptr);
	*out = (T *)ptr;
	return QPDE_OK;
}

// Uploads a [rows][cols] host array with row stride `stride` (0 = one shared
// row) into a dense device array; returns the device stride through dstride.
int upload_rows(qpde_bs1d_plan *p, const double *host, long long stride,
		int rows, int cols, const double **dev, long long *dstride) {
	double *d = nullptr;
	if(stride == 0) {
		int rc = dev_alloc(p, &d, (size_t)cols);
		if(rc) return rc;
		QPDE_CUDA(p->h, cudaMemcpyAsync(d, host, sizeof(double) * cols,
			cudaMemcpyHostToDevice, p->h->stream));
		*dstride = 0;
	} else {
		int rc = dev_alloc(p, &d, (size_t)rows * cols);
		if(rc) return rc;
		if(stride == cols) {
			QPDE_CUDA(p->h, cudaMemcpyAsync(d, host, sizeof(double) * (size_t)rows * cols,
				cudaMemcpyHostToDevice, p->h->stream));      // dense: one copy, not a pitched one
		} else {
			QPDE_CUDA(p->h, cudaMemcpy2DAsync(d, sizeof(double) * cols, host,
				sizeof(double) * stride, sizeof(double) * cols, rows,
				cudaMemcpyHostToDevice, p->h->stream));
		}
		*dstride = cols;
	}
	*dev = d;
	return QPDE_OK;
}

int upload_vec(qpde_bs1d_plan *p, const double *host, int n, const double **dev) {
	double *d = nullptr;
	int rc = dev_alloc(p, &d, (size_t)n);
	if(rc) return rc;
	QPDE_CUDA(p->h, cudaMemcpyAsync(d, host, sizeof(double) * n,
		cudaMemcpyHostToDevice, p->h->stream));
	*dev = d;
	return QPDE_OK;
}

bool is_generator(int kind) {
	return kind >= QPDE_PAYOFF_PUT && kind <= QPDE_PAYOFF_K_MINUS_S;
}

} // namespace

extern "C" {

int qpde_abi_version(void) { return QPDE_ABI_VERSION; }

const char *qpde_last_error(const qpde_handle *h) {
	if(h) return h->error.c_str();
	std::lock_guard<std::mutex> lock(g_err_mutex);
	static thread_local std::string copy;
	copy = g_last_error;
	return copy.c_str();
}

int qpde_create(int device, qpde_handle **out) {
	if(!out) return fail(nullptr, QPDE_ERR_INVALID, "qpde_create: out is NULL");
	*out = nullptr;
	int count = 0;
	cudaError_t err = cudaGetDeviceCount(&count);
	if(err != cudaSuccess || count <= 0) {
		return fail(nullptr, QPDE_ERR_NO_DEVICE,
			"qpde_create: no CUDA device (%s); libqpde_b200 has no CPU fallback",
			err != cudaSuccess ? cudaGetErrorString(err) : "device count is 0");
	}
	if(device < 0 || device >= count) {
		return fail(nullptr, QPDE_ERR_INVALID, "qpde_create: device %d out of range [0, %d)",
			device, count);
	}
	cudaDeviceProp prop;
	QPDE_CUDA(nullptr, cudaGetDeviceProperties(&prop, device));
	if(prop.major != 10) {
		return fail(nullptr, QPDE_ERR_NO_DEVICE,
			"qpde_create: device %d is sm_%d%d; this library holds sm_100a code only",
			device, prop.major, prop.minor);
	}
	QPDE_CUDA(nullptr, cudaSetDevice(device));
	qpde_handle *h = new qpde_handle();
	h->device = device;
	h->sm_count = prop.multiProcessorCount;
	h->launches = 0;
	cudaError_t serr = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
	if(serr != cudaSuccess) {
		delete h;
		return fail(nullptr, QPDE_ERR_CUDA, "cudaStreamCreate failed: %s", cudaGetErrorString(serr));
	}
	if(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking) != cudaSuccess) {
		cudaStreamDestroy(h->stream);
		delete h;
		return fail(nullptr, QPDE_ERR_CUDA, "cudaStreamCreate failed");
	}
	*out = h;
	return QPDE_OK;
}

int qpde_destroy(qpde_handle *h) {
	if(!h) return QPDE_OK;
	cudaSetDevice(h->device);
	cudaStreamSynchronize(h->stream);
	cudaStreamSynchronize(h->copy_stream);
	h->pool.release();
	comm_destroy(h->comm);
	cudaStreamDestroy(h->copy_stream);
	cudaStreamDestroy(h->stream);
	delete h;
	return QPDE_OK;
}

void *qpde_stream(qpde_handle *h) { return h ? (void *)h->stream : nullptr; }

int qpde_synchronize(qpde_handle *h) {
	if(!h) return fail(nullptr, QPDE_ERR_INVALID, "qpde_synchronize: NULL handle");
	QPDE_CUDA(h, cudaSetDevice(h->device));
	QPDE_CUDA(h, cudaStreamSynchronize(h->stream));
	return QPDE_OK;
}

int64_t qpde_launch_count(const qpde_handle *h) { return h ? h->launches : 0; }
int qpde_sm_count(const qpde_handle *h) { return h ? h->sm_count : 0; }

int qpde_measure_fp64_peak(qpde_handle *h, double *tflops) {
	if(!h || !tflops) return fail(h, QPDE_ERR_INVALID, "qpde_measure_fp64_peak: NULL argument");
	QPDE_CUDA(h, cudaSetDevice(h->device));
	const int rc = measure_fp64_peak(h->sm_count, h->stream, tflops, &h->launches);
	if(rc != QPDE_OK) return fail(h, rc, "qpde_measure_fp64_peak failed");
	return QPDE_OK;
}

int qpde_host_alloc(size_t bytes, void **out) {
	if(!out) return fail(nullptr, QPDE_ERR_INVALID, "qpde_host_alloc: out is NULL");
	QPDE_CUDA(nullptr, cudaHostAlloc(out, bytes, cudaHostAllocDefault));
	return QPDE_OK;
}

int qpde_host_free(void *p) {
	if(p) QPDE_CUDA(nullptr, cudaFreeHost(p));
	return QPDE_OK;
}

// ---- host-only helpers -----------------------------------------------------

int qpde_make_schedule(double T, double dt, const double *event_times,
		int n_events, qpde_step *out, int max_steps) {
	// General (non-uniform) event times: same loop as qpde::Replay with an
	// explicit, descending event list.
	if(!(T > 0.) || !(dt > 0.) || n_events < 0) return QPDE_ERR_INVALID;
	std::vector<double> ev(event_times, event_times + n_events);
	std::sort(ev.begin(), ev.end(), [] (double a, double b) { return a > b; });
	int n = 0;
	size_t e = 0;
	double t = T, step = -1., prev;
	do {
		const double nextEventTime = e < ev.size() ? ev[e] : 0.;
		do {
			prev = step;
			step = dt;
			double tmp = t + -1. * step;
			if(fabs(tmp - nextEventTime) < DBL_EPSILON) {
				tmp = nextEventTime;
			} else if(tmp < nextEventTime) {
				tmp = nextEventTime;
				step = -1. * (nextEventTime - t);
			}
			t = tmp;
			if(out && n < max_steps) {
				out[n].t = t; out[n].dt = step; out[n].same_dt = step == prev;
				out[n].event_after = 0; out[n].user_events = 0; out[n].reserved = 0;
			}
			++n;
		} while(nextEventTime < t + -1. * DBL_EPSILON);
		t = nextEventTime;
		int users = 0;
		while(e < ev.size() && ev[e] == t) { ++e; ++users; }
		if(out && n - 1 < max_steps) { out[n - 1].event_after = 1; out[n - 1].user_events = users; }
	} while(0. < t);
	return n;
}

int qpde_make_schedule_forward(double T, double dt, const double *event_times,
		int n_events, qpde_step *out, int max_steps) {
	// qpde::Replay::next_forward with an explicit, ascending event list
	if(!(T > 0.) || !(dt > 0.) || n_events < 0) return QPDE_ERR_INVALID;
	std::vector<double> ev(event_times, event_times + n_events);
	std::sort(ev.begin(), ev.end());
	int n = 0;
	size_t e = 0;
	double t = 0., step = -1., prev;
	do {
		const double nextEventTime = (e < ev.size() && ev[e] <= T) ? ev[e] : T;
		do {
			prev = step;
			step = dt;
			double tmp = t + 1. * step;
			if(fabs(tmp - nextEventTime) < DBL_EPSILON) {
				tmp = nextEventTime;
			} else if(tmp > nextEventTime) {
				tmp = nextEventTime;
				step = 1. * (nextEventTime - t);
			}
			t = tmp;
			if(out && n < max_steps) {
				out[n].t = t; out[n].dt = step; out[n].same_dt = step == prev;
				out[n].event_after = 0; out[n].user_events = 0; out[n].reserved = 0;
			}
			++n;
		} while(nextEventTime > t + 1. * DBL_EPSILON);
		t = nextEventTime;
		int users = 0;
		while(e < ev.size() && ev[e] == t) { ++e; ++users; }
		if(out && n - 1 < max_steps) { out[n - 1].event_after = 1; out[n - 1].user_events = users; }
	} while(T > t);
	return n;
}

int qpde_refined_size(int n_axis, int refine) {
	if(n_axis < 1 || refine < 0 || refine > 20) return QPDE_ERR_INVALID;
	if(n_axis < 2) return n_axis;
	return (1 << refine) * (n_axis - 1) + 1;
}

// ---- plans -----------------------------------------------------------------

int qpde_bs1d_plan_destroy(qpde_bs1d_plan *p) {
	if(!p) return QPDE_OK;
	cudaSetDevice(p->h->device);
	cudaStreamSynchronize(p->h->stream);
	for(auto &blk : p->allocations) p->h->pool.give(blk.first, blk.second);
	delete p;
	return QPDE_OK;
}

int qpde_bs1d_plan_kernel(const qpde_bs1d_plan *p) { return p ? p->kernel : QPDE_ERR_INVALID; }
int qpde_bs1d_plan_nodes(const qpde_bs1d_plan *p) { return p ? p->b.N : QPDE_ERR_INVALID; }
int qpde_bs1d_plan_tma_staged(const qpde_bs1d_plan *p) { return p ? (p->kernel == QPDE_KERNEL_INTERLEAVED && p->tma.enabled ? 1 : 0) : QPDE_ERR_INVALID; }

static int plan_create_impl(qpde_handle *h, const qpde_bs1d_batch *in,
		int outputs, qpde_bs1d_plan **out) {
	if(!h) return fail(nullptr, QPDE_ERR_INVALID, "plan_create: NULL handle");
	if(!in || !out) return fail(h, QPDE_ERR_INVALID, "plan_create: NULL argument");
	*out = nullptr;
	if(in->abi_version != QPDE_ABI_VERSION) {
		return fail(h, QPDE_ERR_INVALID, "batch.abi_version %d != %d", in->abi_version, QPDE_ABI_VERSION);
	}
	const int C = in->n_contracts;
	if(C < 1) return fail(h, QPDE_ERR_INVALID, "n_contracts must be >= 1");
	if(!in->axis || in->n_axis < 3) return fail(h, QPDE_ERR_INVALID, "axis must hold >= 3 ticks");
	if(in->refine < 0 || in->refine > 16) return fail(h, QPDE_ERR_INVALID, "refine out of range");
	if(in->axis_stride != 0 && in->axis_stride < in->n_axis) return fail(h, QPDE_ERR_INVALID, "axis_stride < n_axis");
	if(!in->r || !in->sigma || !in->q || !in->T || !in->dt) {
		return fail(h, QPDE_ERR_INVALID, "r, sigma, q, T and dt are required");
	}
	if(in->scheme < QPDE_SCHEME_RANNACHER || in->scheme > QPDE_SCHEME_BDF6) return fail(h, QPDE_ERR_INVALID, "unknown scheme");
	if(in->vol_model < QPDE_VOL_CONST || in->vol_model > QPDE_VOL_NODES) return fail(h, QPDE_ERR_INVALID, "unknown vol_model");
	if(in->vol_model == QPDE_VOL_NODES && !in->sigma_nodes) return fail(h, QPDE_ERR_INVALID, "sigma_nodes required");
	if(in->payoff_kind == QPDE_PAYOFF_ARRAY ? !in->payoff : !is_generator(in->payoff_kind)) {
		return fail(h, QPDE_ERR_INVALID, "payoff: array missing or unknown generator");
	}
	const bool need_strike = is_generator(in->payoff_kind)
		|| (in->american && is_generator(in->obstacle_kind))
		|| in->n_exercise > 0 || !in->spot;
	if(need_strike && !in->strike) return fail(h, QPDE_ERR_INVALID, "strike required");
	if(in->american) {
		if(in->obstacle_kind == QPDE_PAYOFF_ARRAY ? !in->obstacle : !is_generator(in->obstacle_kind)) {
			return fail(h, QPDE_ERR_INVALID, "obstacle: array missing or unknown generator");
		}
		if(!(in->tolerance > 0.) || !(in->scale > 0.) || !(in->penalty_tolerance > 0.)) {
			return fail(h, QPDE_ERR_INVALID, "tolerance, scale and penalty_tolerance must be > 0");
		}
		if(in->max_inner < 1) return fail(h, QPDE_ERR_INVALID, "max_inner must be >= 1");
	}
	if(in->n_controls != 0) {
		if(in->n_controls < 2 || in->n_controls > 8) return fail(h, QPDE_ERR_INVALID, "n_controls must be 0 or 2 .. 8");
		if(!in->controls) return fail(h, QPDE_ERR_INVALID, "controls required");
		if(in->controls_stride != 0 && in->controls_stride < in->n_controls) return fail(h, QPDE_ERR_INVALID, "controls_stride < n_controls");
		if(in->american || in->n_exercise > 0) return fail(h, QPDE_ERR_UNSUPPORTED, "policy iteration with a penalty constraint or events is not supported");
		if(!(in->tolerance > 0.) || !(in->scale > 0.) || in->max_inner < 1) {
			return fail(h, QPDE_ERR_INVALID, "policy iteration: tolerance, scale > 0 and max_inner >= 1 required");
		}
	}
	if(in->jump_lambda) {
		if(!in->jump_kappa || !in->jump_x0 || !in->jump_dx || !in->jump_weights) {
			return fail(h, QPDE_ERR_INVALID, "jump diffusion: kappa, x0, dx and weights are required (qpde_jump_setup)");
		}
		if(in->n_fft != qpde_jump_nfft(qpde_refined_size(in->n_axis, in->refine))) {
			return fail(h, QPDE_ERR_INVALID, "jump diffusion: n_fft must be qpde_jump_nfft(N)");
		}
		if(in->jump_weights_stride != 0 && in->jump_weights_stride < in->n_fft) return fail(h, QPDE_ERR_INVALID, "jump_weights_stride < n_fft");
		if(in->american || in->n_controls != 0 || in->n_exercise > 0) {
			return fail(h, QPDE_ERR_UNSUPPORTED, "jump diffusion with a penalty constraint, controls or events is not supported");
		}
		if(in->scheme != QPDE_SCHEME_BDF1 && in->scheme != QPDE_SCHEME_BDF2) {
			return fail(h, QPDE_ERR_UNSUPPORTED, "jump diffusion needs scheme BDF1 or BDF2");
		}
	}
	if(in->variable_target) {
		if(in->max_steps < 1) return fail(h, QPDE_ERR_INVALID, "variable time steps need max_steps > 0");
		if(!(in->variable_scale > 0.)) return fail(h, QPDE_ERR_INVALID, "variable_scale must be > 0");
		if(in->n_exercise > 0 || in->jump_lambda) return fail(h, QPDE_ERR_UNSUPPORTED, "variable time steps with events or a jump term are not supported");
		for(int c = 0; c < C; ++c) if(!(in->variable_target[c] > 0.)) return fail(h, QPDE_ERR_INVALID, "contract %d: variable_target must be > 0", c);
	}
	if(in->forward != 0 && in->forward != 1) return fail(h, QPDE_ERR_INVALID, "forward must be 0 or 1");
	if(in->forward && in->variable_target) return fail(h, QPDE_ERR_UNSUPPORTED, "forward time with variable time steps is not supported");
	if(in->n_exercise < 0) return fail(h, QPDE_ERR_INVALID, "n_exercise < 0");
	if(in->n_exercise > 0 && in->exercise_kind != QPDE_EXERCISE_NONE && !is_generator(in->exercise_kind)) {
		return fail(h, QPDE_ERR_INVALID, "exercise_kind must be a generator or QPDE_EXERCISE_NONE");
	}
	if(in->event_times) {
		if(in->n_exercise < 1) return fail(h, QPDE_ERR_INVALID, "event_times needs n_exercise >= 1");
		if(in->variable_target) return fail(h, QPDE_ERR_UNSUPPORTED, "explicit event times with variable steps are not supported");
		double tmax = 0.;
		for(int c = 0; c < C; ++c) tmax = in->T[c] > tmax ? in->T[c] : tmax;
		for(int e = 0; e < in->n_exercise; ++e) {
			if(!(in->event_times[e] >= 0.) || !(in->event_times[e] <= tmax)) return fail(h, QPDE_ERR_INVALID, "event_times[%d] outside [0, T]", e);
		}
		for(int c = 0; c < C; ++c) if(in->T[c] != in->T[0]) return fail(h, QPDE_ERR_INVALID, "explicit event times need one expiry for the whole batch");
	}
	if(in->source_table || in->n_source != 0) {
		if(!in->source_table || in->n_source < 1) return fail(h, QPDE_ERR_INVALID, "source_table / n_source");
		if(in->scheme != QPDE_SCHEME_BDF1 && in->scheme != QPDE_SCHEME_BDF2) return fail(h, QPDE_ERR_UNSUPPORTED, "an operator source needs scheme BDF1 or BDF2");
		if(in->american || in->n_controls != 0 || in->jump_lambda || in->variable_target) {
			return fail(h, QPDE_ERR_UNSUPPORTED, "an operator source with a penalty constraint, controls, a jump term or variable steps is not supported");
		}
	}
	if(in->event_program_length != 0 || in->event_program) {
		if(in->n_exercise < 1) return fail(h, QPDE_ERR_INVALID, "an event program needs n_exercise >= 1 event times");
		if(in->exercise_kind != QPDE_EXERCISE_NONE || in->dividend_kind != QPDE_DIVIDEND_NONE) {
			return fail(h, QPDE_ERR_INVALID, "an event program takes the place of the tagged transforms: exercise_kind must be QPDE_EXERCISE_NONE and dividend_kind QPDE_DIVIDEND_NONE");
		}
		if(in->n_event_consts < 0 || in->n_event_params < 0 || (in->n_event_consts > 0 && !in->event_consts)
				|| (in->n_event_params > 0 && !in->event_params)) return fail(h, QPDE_ERR_INVALID, "event program: constants / parameters missing");
		if(in->event_params_stride != 0 && in->event_params_stride < (long long)in->n_exercise * in->n_event_params) {
			return fail(h, QPDE_ERR_INVALID, "event_params_stride < n_exercise * n_event_params");
		}
		const int bad = event_program_check(in->event_program, in->event_program_length, in->n_event_consts, in->n_event_params);
		if(bad) return fail(h, QPDE_ERR_INVALID, "event program is malformed (check %d: unknown op, operand out of range, or stack depth outside 1 .. %d)", bad, QPDE_EV_MAX_STACK);
		if(in->n_controls != 0 || in->jump_lambda) return fail(h, QPDE_ERR_UNSUPPORTED, "event programs with policy iteration or a jump term are not supported");
	}
	if(in->dividend_kind != QPDE_DIVIDEND_NONE) {
		if(in->dividend_kind != QPDE_DIVIDEND_CASH && in->dividend_kind != QPDE_DIVIDEND_PROPORTIONAL) return fail(h, QPDE_ERR_INVALID, "unknown dividend_kind");
		if(!in->dividend || in->n_exercise < 1) return fail(h, QPDE_ERR_INVALID, "dividends need the amounts and n_exercise >= 1 event times");
		if(in->n_controls != 0) return fail(h, QPDE_ERR_UNSUPPORTED, "policy iteration with events is not supported");
		for(int c = 0; c < C; ++c) {
			const double D = in->dividend[c];
			if(!(D >= 0.) || (in->dividend_kind == QPDE_DIVIDEND_PROPORTIONAL && !(D < 1.))) {
				return fail(h, QPDE_ERR_INVALID, "contract %d: dividend must be >= 0 (and < 1 when proportional)", c);
			}
		}
	}
	for(int c = 0; c < C; ++c) {
		if(!(in->T[c] > 0.) || !(in->dt[c] > DBL_EPSILON) || in->dt[c] > in->T[c]) {
			return fail(h, QPDE_ERR_INVALID, "contract %d: need 0 < dt <= T", c);
		}
	}

	QPDE_CUDA(h, cudaSetDevice(h->device));
	qpde_bs1d_plan *p = new qpde_bs1d_plan();
	p->h = h;
	p->outputs = outputs;
	memset(&p->b, 0, sizeof(p->b));
	memset(&p->out, 0, sizeof(p->out));
	memset(&p->w, 0, sizeof(p->w));
	memset(&p->jw, 0, sizeof(p->jw));
	p->jump = in->jump_lambda != nullptr;
	DeviceBatch &b = p->b;
	b.C = C;
	b.n_axis = in->n_axis; b.refine = in->refine;
	b.N = qpde_refined_size(in->n_axis, in->refine);
	const int N = b.N;
	b.vol_model = in->vol_model;
	b.scheme = in->scheme;
	b.forward = in->forward ? 1 : 0;
	b.n_exercise = in->n_exercise; b.exercise_kind = in->exercise_kind;
	b.dividend_kind = in->dividend_kind; b.dividend_first = in->dividend_first ? 1 : 0;
	b.payoff_kind = in->payoff_kind;
	b.obstacle_kind = in->american ? in->obstacle_kind : 0;
	b.american = in->american ? 1 : 0;
	b.max_inner = in->max_inner;
	b.tolerance = in->tolerance; b.scale = in->scale; b.penalty_tolerance = in->penalty_tolerance;
	b.pen = in->american ? 1. / in->penalty_tolerance : 0.;
	b.n_controls = in->n_controls; b.policy_max = in->policy_max ? 1 : 0;

	// Row length of the per-step outputs.  Constant steps: at least the bound of the replayed loop (a
	// smaller caller value would put inner[] out of bounds); variable steps: the caller's capacity.
	int max_steps = in->max_steps > 0 ? in->max_steps : 0;
	if(!in->variable_target) {
		for(int c = 0; c < C; ++c) {
			const int bound = replay_step_bound(in->T[c], in->dt[c], in->n_exercise);
			if(bound < 0) {
				delete p;
				return fail(h, QPDE_ERR_INVALID, "contract %d: T / dt is too large (more than 2^30 steps)", c);
			}
			if(bound > max_steps) max_steps = bound;
		}
	}
	b.max_steps = max_steps;

	int rc = QPDE_OK;
#define QPDE_TRY(expr) do { rc = (expr); if(rc != QPDE_OK) { qpde_bs1d_plan_destroy(p); return rc; } } while(0)
	QPDE_TRY(upload_rows(p, in->axis, in->axis_stride, C, in->n_axis, &b.axis, &b.axis_stride));
	QPDE_TRY(upload_vec(p, in->r, C, &b.r));
	QPDE_TRY(upload_vec(p, in->sigma, C, &b.sigma));
	QPDE_TRY(upload_vec(p, in->q, C, &b.q));
	QPDE_TRY(upload_vec(p, in->T, C, &b.T));
	QPDE_TRY(upload_vec(p, in->dt, C, &b.dt));
	if(in->strike) QPDE_TRY(upload_vec(p, in->strike, C, &b.strike));
	else           QPDE_TRY(upload_vec(p, in->spot, C, &b.strike));
	if(in->spot) QPDE_TRY(upload_vec(p, in->spot, C, &b.spot));
	if(in->variable_target) { QPDE_TRY(upload_vec(p, in->variable_target, C, &b.variable_target)); b.variable_scale = in->variable_scale; }
	if(in->dividend_kind != QPDE_DIVIDEND_NONE) QPDE_TRY(upload_vec(p, in->dividend, C, &b.dividend));
	if(in->event_times) {
		std::vector<double> sorted(in->event_times, in->event_times + in->n_exercise);
		std::sort(sorted.begin(), sorted.end());
		QPDE_TRY(upload_vec(p, sorted.data(), in->n_exercise, &b.event_times));
		if(cudaStreamSynchronize(h->stream) != cudaSuccess) {      // `sorted` goes away
			qpde_bs1d_plan_destroy(p);
			return fail(h, QPDE_ERR_CUDA, "event_times upload failed");
		}
	}
	if(in->source_table) { QPDE_TRY(upload_vec(p, in->source_table, in->n_source, &b.source_table)); b.n_source = in->n_source; }
	if(in->event_program) {
		// program words (as doubles' worth of ints), constants and the per-event parameter tables
		int *code = nullptr;
		QPDE_TRY(dev_alloc(p, &code, (size_t)in->event_program_length));
		if(cudaMemcpyAsync(code, in->event_program, sizeof(int) * in->event_program_length, cudaMemcpyHostToDevice, h->stream) != cudaSuccess) {
			qpde_bs1d_plan_destroy(p);
			return fail(h, QPDE_ERR_CUDA, "event program upload failed");
		}
		b.event_program = code;
		const double zero = 0.;
		QPDE_TRY(upload_vec(p, in->n_event_consts > 0 ? in->event_consts : &zero, in->n_event_consts > 0 ? in->n_event_consts : 1, &b.event_consts));
		const int row = in->n_exercise * in->n_event_params;
		if(row > 0) QPDE_TRY(upload_rows(p, in->event_params, in->event_params_stride, C, row, &b.event_params, &b.event_params_stride));
		else { QPDE_TRY(upload_vec(p, &zero, 1, &b.event_params)); b.event_params_stride = 0; }
		b.n_event_params = in->n_event_params;
	}
	if(p->jump) {
		b.n_fft = in->n_fft;
		QPDE_TRY(upload_vec(p, in->jump_lambda, C, &b.jump_lambda));
		QPDE_TRY(upload_vec(p, in->jump_kappa, C, &b.jump_kappa));
		QPDE_TRY(upload_vec(p, in->jump_x0, C, &b.jump_x0));
		QPDE_TRY(upload_vec(p, in->jump_dx, C, &b.jump_dx));
		QPDE_TRY(upload_rows(p, in->jump_weights, in->jump_weights_stride, C, in->n_fft,
			&b.jump_weights, &b.jump_weights_stride));
	}
	if(in->n_controls > 0) {
		QPDE_TRY(upload_rows(p, in->controls, in->controls_stride, C, in->n_controls, &b.controls, &b.controls_stride));
	}
	if(in->vol_model == QPDE_VOL_NODES) {
		QPDE_TRY(upload_rows(p, in->sigma_nodes, in->sigma_nodes_stride ? in->sigma_nodes_stride : 0,
			C, N, &b.sigma_nodes, &b.sigma_nodes_stride));
	}
	if(in->payoff_kind == QPDE_PAYOFF_ARRAY) {
		QPDE_TRY(upload_rows(p, in->payoff, in->payoff_stride, C, N, &b.payoff, &b.payoff_stride));
	}
	if(in->american && in->obstacle_kind == QPDE_PAYOFF_ARRAY) {
		QPDE_TRY(upload_rows(p, in->obstacle, in->obstacle_stride, C, N, &b.obstacle, &b.obstacle_stride));
	}

	// outputs
	DeviceResult &o = p->out;
	if(outputs & QPDE_OUT_V)      { QPDE_TRY(dev_alloc(p, &o.V, (size_t)C * N)); o.V_stride = N; }
	if(outputs & QPDE_OUT_S)      { QPDE_TRY(dev_alloc(p, &o.S, (size_t)C * N)); o.S_stride = N; }
	if(outputs & QPDE_OUT_VALUE)  QPDE_TRY(dev_alloc(p, &o.value, (size_t)C));
	if(outputs & QPDE_OUT_STEPS)  QPDE_TRY(dev_alloc(p, &o.n_steps, (size_t)C));
	if(outputs & QPDE_OUT_INNER_TOTAL) QPDE_TRY(dev_alloc(p, &o.inner_total, (size_t)C));
	if(outputs & QPDE_OUT_COMPUTED) QPDE_TRY(dev_alloc(p, &o.computed, (size_t)C));
	if(outputs & QPDE_OUT_INNER)  { QPDE_TRY(dev_alloc(p, &o.inner, (size_t)C * max_steps)); o.inner_stride = max_steps; }
	if(outputs & QPDE_OUT_ACTIVE) { QPDE_TRY(dev_alloc(p, &o.active, (size_t)C * N)); o.active_stride = N; }
	if(outputs & QPDE_OUT_STATUS) QPDE_TRY(dev_alloc(p, &o.status, (size_t)C));
	if(o.inner && cudaMemsetAsync(o.inner, 0, (size_t)C * max_steps, h->stream) != cudaSuccess) {
		qpde_bs1d_plan_destroy(p);
		return fail(h, QPDE_ERR_CUDA, "cudaMemsetAsync(inner) failed: %s", cudaGetErrorString(cudaGetLastError()));
	}

	// kernel choice
	int kernel = in->kernel;
	if(p->jump) {
		// the jump-diffusion sweep exists in the RESIDENT family only
		if(kernel == QPDE_KERNEL_INTERLEAVED || !jump_supported(b)) {
			qpde_bs1d_plan_destroy(p);
			return fail(h, QPDE_ERR_UNSUPPORTED, "jump diffusion: N = %d, n_fft = %d is not supported by the RESIDENT family", N, b.n_fft);
		}
		kernel = QPDE_KERNEL_RESIDENT;
		QPDE_TRY(dev_alloc(p, &p->jw.g_idx, (size_t)C * b.n_fft));
		QPDE_TRY(dev_alloc(p, &p->jw.g_w, (size_t)C * b.n_fft));
		QPDE_TRY(dev_alloc(p, &p->jw.n_idx, (size_t)C * N));
		QPDE_TRY(dev_alloc(p, &p->jw.n_w, (size_t)C * N));
	}
	if(kernel == QPDE_KERNEL_AUTO) {
		kernel = resident_supported(b) ? QPDE_KERNEL_RESIDENT : QPDE_KERNEL_INTERLEAVED;
	}
	if(kernel == QPDE_KERNEL_RESIDENT && !p->jump && !resident_supported(b)) {
		qpde_bs1d_plan_destroy(p);
		return fail(h, QPDE_ERR_UNSUPPORTED, "the RESIDENT kernel does not support N = %d with this configuration", N);
	}
	if(kernel != QPDE_KERNEL_RESIDENT && kernel != QPDE_KERNEL_INTERLEAVED) {
		qpde_bs1d_plan_destroy(p);
		return fail(h, QPDE_ERR_INVALID, "unknown kernel %d", kernel);
	}
	p->kernel = kernel;

	if(kernel == QPDE_KERNEL_INTERLEAVED) {
		InterleavedWork &w = p->w;
		w.ld = ((long long)C + 31) / 32 * 32;                 // a warp of contracts per 256-byte row segment
		const size_t NC = (size_t)N * (size_t)w.ld;
		const size_t sets = b.n_controls > 0 ? (size_t)b.n_controls : 1;   // operator rows per control value
		// (policy iteration under a theta scheme re-forms the right-hand side from V^n at every control update)
		const bool need_xprev = bdf_order(b.scheme) >= 2 || b.variable_target || (b.n_controls > 0 && bdf_order(b.scheme) == 0);
		const size_t hist_slots = bdf_order(b.scheme) > 2 ? (size_t)bdf_order(b.scheme) - 1 : 1;   // iterand(1) .. iterand(order - 1)
		const bool need_evobst = b.n_exercise > 0 && b.exercise_kind != QPDE_EXERCISE_NONE;
		p->tma.enabled = false;
		p->wblock = nullptr;
		const char *no_tma = getenv("QPDE_INTERLEAVED_NO_TMA");  // tuning / A-B switch: the plain-load sweep
		if(interleaved_tma_supported(b) && !(no_tma && no_tma[0] == '1')) {
			// one block, arrays stacked: tensor maps over [array][rows][32] serve every array
			const size_t arrays = 8 + (need_xprev ? 1 : 0) + (b.american ? 1 : 0) + (need_evobst ? 1 : 0);
			QPDE_TRY(dev_alloc(p, &p->wblock, NC * arrays));
			double *at = p->wblock;
			auto take = [&](double **ptr) { *ptr = at; at += NC; };
			// arrays that a pass reads together are adjacent: (lo, di, up), (obstacle, x), (cp, dp) each
			// arrive with one TMA box
			take(&w.S); take(&w.lo); take(&w.di); take(&w.up); take(&w.lb);
			if(b.american) take(&w.obst);
			take(&w.x); take(&w.cp); take(&w.dp);
			if(need_xprev) take(&w.xprev); else w.xprev = w.x;
			if(need_evobst) take(&w.evobst);
			// lanes past C in the last warp of contracts are loaded (never stored): keep them finite
			if(cudaMemsetAsync(p->wblock, 0, NC * arrays * sizeof(double), h->stream) != cudaSuccess) {
				qpde_bs1d_plan_destroy(p);
				return fail(h, QPDE_ERR_CUDA, "cudaMemsetAsync(workspace) failed: %s", cudaGetErrorString(cudaGetLastError()));
			}
			std::string err;
			if(!interleaved_tma_prepare(&p->tma, p->wblock, (long long)(NC / QPDE_WARP_ROW), (int)arrays, &err)) {
				qpde_bs1d_plan_destroy(p);
				return fail(h, QPDE_ERR_CUDA, "INTERLEAVED (TMA): %s", err.c_str());
			}
		} else {
			QPDE_TRY(dev_alloc(p, &w.S, NC));  QPDE_TRY(dev_alloc(p, &w.lo, NC * sets));
			QPDE_TRY(dev_alloc(p, &w.di, NC * sets)); QPDE_TRY(dev_alloc(p, &w.up, NC * sets));
			QPDE_TRY(dev_alloc(p, &w.x, NC));  QPDE_TRY(dev_alloc(p, &w.lb, NC));
			QPDE_TRY(dev_alloc(p, &w.cp, NC)); QPDE_TRY(dev_alloc(p, &w.dp, NC));
			if(need_xprev) QPDE_TRY(dev_alloc(p, &w.xprev, NC * hist_slots));
			else w.xprev = w.x;
			if(b.american) QPDE_TRY(dev_alloc(p, &w.obst, NC));
			if(need_evobst) QPDE_TRY(dev_alloc(p, &w.evobst, NC));
		}
	}
#undef QPDE_TRY
	*out = p;
	return QPDE_OK;
}

int qpde_bs1d_plan_create(qpde_handle *h, const qpde_bs1d_batch *batch, qpde_bs1d_plan **out) {
	const int outputs = batch && batch->outputs ? batch->outputs : QPDE_OUT_ALL;
	return plan_create_impl(h, batch, outputs, out);
}

int qpde_bs1d_plan_run(qpde_bs1d_plan *p) {
	if(!p) return fail(nullptr, QPDE_ERR_INVALID, "plan_run: NULL plan");
	qpde_handle *h = p->h;
	QPDE_CUDA(h, cudaSetDevice(h->device));
	if(p->jump) {
		int rc = launch_jump(p->b, p->jw, p->out, h->stream, &h->launches);
		if(rc != QPDE_OK) return fail(h, rc, "launch_jump failed");
	} else if(p->kernel == QPDE_KERNEL_INTERLEAVED) {
		if(p->tma.enabled) launch_interleaved_tma(p->b, p->w, p->tma, p->wblock, p->out, h->stream, &h->launches);
		else launch_interleaved(p->b, p->w, p->out, h->stream, &h->launches);
	} else {
		int rc = launch_resident(p->b, p->out, h->stream, h->sm_count, &h->launches);
		if(rc != QPDE_OK) return fail(h, rc, "launch_resident failed");
	}
	QPDE_CUDA(h, cudaGetLastError());
	return QPDE_OK;
}

// Copies the outputs of contracts [c0, c1) to the host buffers on `stream` (no synchronisation).
static int fetch_range(qpde_bs1d_plan *p, qpde_bs1d_result *r, int c0, int c1, cudaStream_t stream) {
	qpde_handle *h = p->h;
	const int N = p->b.N;
	const size_t n = (size_t)(c1 - c0);
	const DeviceResult &o = p->out;
	if(r->V) {
		const long long hs = r->V_stride ? r->V_stride : N;
		if(hs < N) return fail(h, QPDE_ERR_INVALID, "V_stride < N");
		if(hs == N && o.V_stride == N) {
			QPDE_CUDA(h, cudaMemcpyAsync(r->V + (size_t)c0 * N, o.V + (size_t)c0 * N, sizeof(double) * (size_t)N * n, cudaMemcpyDeviceToHost, stream));
		} else
		QPDE_CUDA(h, cudaMemcpy2DAsync(r->V + (size_t)c0 * hs, sizeof(double) * hs, o.V + (size_t)c0 * o.V_stride, sizeof(double) * o.V_stride,
			sizeof(double) * N, n, cudaMemcpyDeviceToHost, stream));
	}
	if(r->S) {
		const long long hs = r->S_stride ? r->S_stride : N;
		if(hs < N) return fail(h, QPDE_ERR_INVALID, "S_stride < N");
		QPDE_CUDA(h, cudaMemcpy2DAsync(r->S + (size_t)c0 * hs, sizeof(double) * hs, o.S + (size_t)c0 * o.S_stride, sizeof(double) * o.S_stride,
			sizeof(double) * N, n, cudaMemcpyDeviceToHost, stream));
	}
	if(r->value) QPDE_CUDA(h, cudaMemcpyAsync(r->value + c0, o.value + c0, sizeof(double) * n, cudaMemcpyDeviceToHost, stream));
	if(r->n_steps) QPDE_CUDA(h, cudaMemcpyAsync(r->n_steps + c0, o.n_steps + c0, sizeof(int) * n, cudaMemcpyDeviceToHost, stream));
	if(r->inner_total) QPDE_CUDA(h, cudaMemcpyAsync(r->inner_total + c0, o.inner_total + c0, sizeof(int) * n, cudaMemcpyDeviceToHost, stream));
	if(r->inner) {
		const long long hs = r->inner_stride ? r->inner_stride : p->b.max_steps;
		const long long width = hs < o.inner_stride ? hs : o.inner_stride;
		QPDE_CUDA(h, cudaMemcpy2DAsync(r->inner + (size_t)c0 * hs, (size_t)hs, o.inner + (size_t)c0 * o.inner_stride, (size_t)o.inner_stride,
			(size_t)width, n, cudaMemcpyDeviceToHost, stream));
	}
	if(r->active) {
		const long long hs = r->active_stride ? r->active_stride : N;
		if(hs < N) return fail(h, QPDE_ERR_INVALID, "active_stride < N");
		QPDE_CUDA(h, cudaMemcpy2DAsync(r->active + (size_t)c0 * hs, (size_t)hs, o.active + (size_t)c0 * o.active_stride, (size_t)o.active_stride,
			(size_t)N, n, cudaMemcpyDeviceToHost, stream));
	}
	if(r->status) QPDE_CUDA(h, cudaMemcpyAsync(r->status + c0, o.status + c0, sizeof(int) * n, cudaMemcpyDeviceToHost, stream));
	if(r->solves_computed) QPDE_CUDA(h, cudaMemcpyAsync(r->solves_computed + c0, o.computed + c0, sizeof(int) * n, cudaMemcpyDeviceToHost, stream));
	return QPDE_OK;
}

int qpde_bs1d_plan_fetch(qpde_bs1d_plan *p, qpde_bs1d_result *r) {
	if(!p || !r) return fail(nullptr, QPDE_ERR_INVALID, "plan_fetch: NULL argument");
	qpde_handle *h = p->h;
	const DeviceResult &o = p->out;
	QPDE_CUDA(h, cudaSetDevice(h->device));
#define QPDE_NEED(hostptr, devptr, name) \
	if((hostptr) && !(devptr)) return fail(h, QPDE_ERR_INVALID, "output '%s' was not enabled in batch.outputs", name)
	QPDE_NEED(r->V, o.V, "V"); QPDE_NEED(r->S, o.S, "S"); QPDE_NEED(r->value, o.value, "value");
	QPDE_NEED(r->n_steps, o.n_steps, "n_steps"); QPDE_NEED(r->inner_total, o.inner_total, "inner_total");
	QPDE_NEED(r->inner, o.inner, "inner"); QPDE_NEED(r->active, o.active, "active");
	QPDE_NEED(r->status, o.status, "status"); QPDE_NEED(r->solves_computed, o.computed, "solves_computed");
#undef QPDE_NEED
	const int rc = fetch_range(p, r, 0, p->b.C, h->stream);
	if(rc != QPDE_OK) return rc;
	QPDE_CUDA(h, cudaStreamSynchronize(h->stream));
	return QPDE_OK;
}

// The one-shot call on a large RESIDENT batch: the contracts are swept in a few ranges, and the
// results of a finished range travel to the host (second stream, ordered by an event) while the
// next range is being swept — end to end the call costs the sweep plus the copy of the last range.
static int run_and_fetch_ranges(qpde_bs1d_plan *p, qpde_bs1d_result *r) {
	qpde_handle *h = p->h;
	const int C = p->b.C;
	// whole waves per range: two resident contracts per SM at most
	const int wave = 2 * h->sm_count;
	int per = ((C + 3) / 4 + wave - 1) / wave * wave;
	if(per < 8 * wave) per = 8 * wave;
	const int ranges = (C + per - 1) / per;
	std::vector<cudaEvent_t> done((size_t)ranges);
	int rc = QPDE_OK;
	int made = 0;
	for(; made < ranges; ++made) if(cudaEventCreateWithFlags(&done[made], cudaEventDisableTiming) != cudaSuccess) { rc = QPDE_ERR_CUDA; break; }
	for(int k = 0; rc == QPDE_OK && k < ranges; ++k) {
		DeviceBatch b = p->b;
		b.c0 = k * per;
		b.C = C - b.c0 < per ? C - b.c0 : per;
		rc = launch_resident(b, p->out, h->stream, h->sm_count, &h->launches);
		if(rc == QPDE_OK && cudaEventRecord(done[k], h->stream) != cudaSuccess) rc = QPDE_ERR_CUDA;
	}
	for(int k = 0; rc == QPDE_OK && k < ranges; ++k) {
		const int c0 = k * per, c1 = c0 + per < C ? c0 + per : C;
		if(cudaStreamWaitEvent(h->copy_stream, done[k], 0) != cudaSuccess) { rc = QPDE_ERR_CUDA; break; }
		rc = fetch_range(p, r, c0, c1, h->copy_stream);
	}
	const cudaError_t e1 = cudaStreamSynchronize(h->stream), e2 = cudaStreamSynchronize(h->copy_stream);
	for(int k = 0; k < made; ++k) cudaEventDestroy(done[k]);
	if(rc != QPDE_OK) return rc == QPDE_ERR_CUDA ? fail(h, rc, "ranged sweep failed: %s", cudaGetErrorString(cudaGetLastError())) : rc;
	if(e1 != cudaSuccess || e2 != cudaSuccess) return fail(h, QPDE_ERR_CUDA, "ranged sweep failed: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
	return QPDE_OK;
}

} // extern "C" (reopened below)

// One validation for the single-GPU solve and the sharded entry points.
static int gmwb_validate(qpde_handle *h, const qpde_gmwb2d *g) {
	if(g->abi_version != QPDE_ABI_VERSION) return fail(h, QPDE_ERR_INVALID, "problem.abi_version %d != %d", g->abi_version, QPDE_ABI_VERSION);
	if(!g->w_axis || !g->a_axis || g->n_w_axis < 3 || g->n_a_axis < 3) return fail(h, QPDE_ERR_INVALID, "gmwb: axes need >= 3 ticks");
	if(!g->controls || g->n_controls < 1 || g->n_controls > 16) return fail(h, QPDE_ERR_INVALID, "gmwb: 1 .. 16 stochastic controls");
	if(!g->impulse_axis || g->n_impulse_axis < 2) return fail(h, QPDE_ERR_INVALID, "gmwb: impulse axis needs >= 2 ticks");
	if(g->refine < 0 || g->refine > 8) return fail(h, QPDE_ERR_INVALID, "gmwb: refine out of range");
	if(!(g->T > 0.) || g->timesteps < 1 || g->max_inner < 1 || g->max_sweeps < 1) return fail(h, QPDE_ERR_INVALID, "gmwb: T, timesteps, max_inner, max_sweeps must be positive");
	if(!(g->scaling_factor > 0.) || !(g->tolerance > 0.)) return fail(h, QPDE_ERR_INVALID, "gmwb: scaling_factor and tolerance must be > 0");
	if(g->w_axis[0] != 0. || g->a_axis[0] != 0.) return fail(h, QPDE_ERR_INVALID, "gmwb: both axes start at 0 (no left boundary routine, as in the reference model)");
	return QPDE_OK;
}

extern "C" {

int qpde_gmwb2d_nodes(const qpde_gmwb2d *g, int *nodes_w, int *nodes_a) {
	if(!g || !nodes_w || !nodes_a) return QPDE_ERR_INVALID;
	*nodes_w = qpde_refined_size(g->n_w_axis, g->refine);
	*nodes_a = qpde_refined_size(g->n_a_axis, g->refine);
	return *nodes_w > 0 && *nodes_a > 0 ? QPDE_OK : QPDE_ERR_INVALID;
}

// ---- 1-D impulse-control HJBQVI (qpde_hjbqvi1d.cu) ------------------------------------------

int qpde_hjbqvi1d_nodes(const qpde_hjbqvi1d_batch *g, int *nodes, int *n_controls, int *n_impulses) {
	if(!g || !nodes) return QPDE_ERR_INVALID;
	*nodes = qpde_refined_size(g->n_x_axis, g->refine);
	if(n_controls) *n_controls = qpde_refined_size(g->n_control_axis, g->refine);
	if(n_impulses) *n_impulses = qpde_refined_size(g->n_impulse_axis, g->refine);
	return *nodes > 0 ? QPDE_OK : QPDE_ERR_INVALID;
}

int qpde_hjbqvi1d_solve(qpde_handle *h, const qpde_hjbqvi1d_batch *g, qpde_hjbqvi1d_result *res) {
	if(!h) return fail(nullptr, QPDE_ERR_INVALID, "qpde_hjbqvi1d_solve: NULL handle");
	if(!g || !res || !res->V) return fail(h, QPDE_ERR_INVALID, "qpde_hjbqvi1d_solve: NULL argument");
	if(g->abi_version != QPDE_ABI_VERSION) return fail(h, QPDE_ERR_INVALID, "batch.abi_version %d != %d", g->abi_version, QPDE_ABI_VERSION);
	if(g->model != QPDE_HJBQVI1D_EXCHANGE_RATE) return fail(h, QPDE_ERR_UNSUPPORTED, "hjbqvi1d: unknown model %d", g->model);
	const int P = g->n_problems;
	if(P < 1) return fail(h, QPDE_ERR_INVALID, "hjbqvi1d: n_problems must be >= 1");
	if(g->refine < 0 || g->refine > 12) return fail(h, QPDE_ERR_INVALID, "hjbqvi1d: refine out of range");
	if(!g->x_axis || g->n_x_axis < 3 || !g->control_axis || g->n_control_axis < 1 || !g->impulse_axis || g->n_impulse_axis < 1) {
		return fail(h, QPDE_ERR_INVALID, "hjbqvi1d: axes missing (x needs >= 3 ticks, the control grids >= 1)");
	}
	if((g->x_axis_stride != 0 && g->x_axis_stride < g->n_x_axis) || (g->control_axis_stride != 0 && g->control_axis_stride < g->n_control_axis)
			|| (g->impulse_axis_stride != 0 && g->impulse_axis_stride < g->n_impulse_axis)) return fail(h, QPDE_ERR_INVALID, "hjbqvi1d: axis stride below the axis length");
	if(!g->params || g->n_params < 7 || (g->params_stride != 0 && g->params_stride < g->n_params)) return fail(h, QPDE_ERR_INVALID, "hjbqvi1d: the exchange-rate model takes 7 parameters per problem");
	if(!(g->T > 0.) || g->timesteps < 1 || g->max_inner < 1 || g->max_sweeps < 1) return fail(h, QPDE_ERR_INVALID, "hjbqvi1d: T, timesteps, max_inner and max_sweeps must be positive");
	if(!(g->scaling_factor > 0.) || !(g->tolerance > 0.)) return fail(h, QPDE_ERR_INVALID, "hjbqvi1d: scaling_factor and tolerance must be > 0");
	const int N = qpde_refined_size(g->n_x_axis, g->refine);
	const int single = g->n_control_axis < 2 ? 0 : g->refine, single_z = g->n_impulse_axis < 2 ? 0 : g->refine;
	const int nq = qpde_refined_size(g->n_control_axis, single), nz = qpde_refined_size(g->n_impulse_axis, single_z);
	if(N > QPDE_HJBQVI1D_MAX_NODES) return fail(h, QPDE_ERR_UNSUPPORTED, "hjbqvi1d: %d nodes, at most %d", N, QPDE_HJBQVI1D_MAX_NODES);
	if(hjbqvi1d_shared_bytes(N, nq, nz) > (size_t)227 * 1024) return fail(h, QPDE_ERR_UNSUPPORTED, "hjbqvi1d: %d + %d control nodes do not fit the shared memory of an SM", nq, nz);
	for(int p = 0; p < P; ++p) {
		const double *ax[3] = { g->x_axis + (size_t)p * g->x_axis_stride, g->control_axis + (size_t)p * g->control_axis_stride, g->impulse_axis + (size_t)p * g->impulse_axis_stride };
		const int len[3] = { g->n_x_axis, g->n_control_axis, g->n_impulse_axis };
		for(int k = 0; k < 3; ++k) for(int i = 1; i < len[k]; ++i) if(!(ax[k][i] > ax[k][i - 1])) return fail(h, QPDE_ERR_INVALID, "hjbqvi1d: problem %d: axis %d is not strictly increasing", p, k);
		const double *pp = g->params + (size_t)p * g->params_stride;
		if(!(pp[1] >= 0.)) return fail(h, QPDE_ERR_INVALID, "hjbqvi1d: problem %d: sigma < 0", p);
	}
	QPDE_CUDA(h, cudaSetDevice(h->device));
	// device copies: inputs in one block, outputs in another
	const size_t nx_in = (g->x_axis_stride ? (size_t)P : 1) * g->n_x_axis, nq_in = (g->control_axis_stride ? (size_t)P : 1) * g->n_control_axis;
	const size_t nz_in = (g->impulse_axis_stride ? (size_t)P : 1) * g->n_impulse_axis, np_in = (g->params_stride ? (size_t)P : 1) * g->n_params;
	std::vector<double> in(nx_in + nq_in + nz_in + np_in);
	auto pack = [&](double *dst, const double *src, long long stride, int len, size_t total) {
		if(stride == 0) { memcpy(dst, src, sizeof(double) * len); return; }
		for(size_t p = 0; p * len < total; ++p) memcpy(dst + p * len, src + p * (size_t)stride, sizeof(double) * len);
	};
	pack(in.data(), g->x_axis, g->x_axis_stride, g->n_x_axis, nx_in);
	pack(in.data() + nx_in, g->control_axis, g->control_axis_stride, g->n_control_axis, nq_in);
	pack(in.data() + nx_in + nq_in, g->impulse_axis, g->impulse_axis_stride, g->n_impulse_axis, nz_in);
	pack(in.data() + nx_in + nq_in + nz_in, g->params, g->params_stride, g->n_params, np_in);
	const size_t PN = (size_t)P * N;
	const size_t out_bytes = sizeof(double) * (4 * PN + 2 * P) + sizeof(int) * 2 * (size_t)P + PN;
	double *d_in = nullptr; unsigned char *d_out = nullptr;
	QPDE_CUDA(h, cudaMalloc((void **)&d_in, sizeof(double) * in.size()));
	if(cudaMalloc((void **)&d_out, out_bytes) != cudaSuccess) { cudaFree(d_in); return fail(h, QPDE_ERR_CUDA, "hjbqvi1d: out of device memory"); }
	int rc = QPDE_OK;
	do {
		if(cudaMemcpyAsync(d_in, in.data(), sizeof(double) * in.size(), cudaMemcpyHostToDevice, h->stream) != cudaSuccess) { rc = QPDE_ERR_CUDA; break; }
		Hjbqvi1dBatch b;
		b.n_problems = P; b.model = g->model; b.N = N; b.nq = nq; b.nz = nz;
		b.refine = g->refine; b.refine_q = single; b.refine_z = single_z;
		b.x_axis = d_in; b.q_axis = d_in + nx_in; b.z_axis = d_in + nx_in + nq_in; b.params = d_in + nx_in + nq_in + nz_in;
		b.x_axis_stride = g->x_axis_stride ? g->n_x_axis : 0; b.q_axis_stride = g->control_axis_stride ? g->n_control_axis : 0;
		b.z_axis_stride = g->impulse_axis_stride ? g->n_impulse_axis : 0; b.params_stride = g->params_stride ? g->n_params : 0;
		b.T = g->T; b.timesteps = g->timesteps; b.max_inner = g->max_inner; b.max_sweeps = g->max_sweeps;
		b.scaling_factor = g->scaling_factor; b.tolerance = g->tolerance;
		Hjbqvi1dResult o;
		double *dd = reinterpret_cast<double *>(d_out);
		o.V = dd; o.X = dd + PN; o.Q = dd + 2 * PN; o.Z = dd + 3 * PN; o.mean_inner = dd + 4 * PN; o.mean_sweeps = dd + 4 * PN + P;
		o.n_steps = reinterpret_cast<int *>(dd + 4 * PN + 2 * P); o.status = o.n_steps + P;
		o.mask = reinterpret_cast<unsigned char *>(o.status + P);
		rc = launch_hjbqvi1d(b, o, h->stream, &h->launches);
		if(rc != QPDE_OK) break;
		auto fetch = [&](void *dst, const void *src, size_t bytes) { return !dst || cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, h->stream) == cudaSuccess; };
		const bool ok = fetch(res->V, o.V, sizeof(double) * PN) && fetch(res->X, o.X, sizeof(double) * PN) && fetch(res->Q, o.Q, sizeof(double) * PN)
			&& fetch(res->Z, o.Z, sizeof(double) * PN) && fetch(res->mask, o.mask, PN) && fetch(res->mean_inner, o.mean_inner, sizeof(double) * P) && fetch(res->mean_sweeps, o.mean_sweeps, sizeof(double) * P)
			&& fetch(res->n_steps, o.n_steps, sizeof(int) * P) && fetch(res->status, o.status, sizeof(int) * P);
		if(!ok || cudaStreamSynchronize(h->stream) != cudaSuccess) rc = QPDE_ERR_CUDA;
	} while(0);
	const cudaError_t last = cudaGetLastError();
	cudaFree(d_in); cudaFree(d_out);
	if(rc == QPDE_ERR_UNSUPPORTED) return fail(h, rc, "hjbqvi1d: this grid / model is not covered by the kernel");
	if(rc != QPDE_OK) return fail(h, rc, "hjbqvi1d: CUDA error: %s", cudaGetErrorString(last));
	return QPDE_OK;
}

// ---- 2-D impulse-control HJBQVI with impulses both ways (qpde_hjbqvi2d.cu) -------------------

int qpde_hjbqvi2d_nodes(const qpde_hjbqvi2d *g, int *nodes_w, int *nodes_b) {
	if(!g || !nodes_w || !nodes_b) return QPDE_ERR_INVALID;
	*nodes_w = qpde_refined_size(g->n_w_axis, g->refine);
	*nodes_b = qpde_refined_size(g->n_b_axis, g->refine);
	return *nodes_w > 0 && *nodes_b > 0 ? QPDE_OK : QPDE_ERR_INVALID;
}

int qpde_hjbqvi2d_solve(qpde_handle *h, const qpde_hjbqvi2d *g, double *V, double *Q, double *Z, unsigned char *mask,
		double *W_out, double *B_out, qpde_hjbqvi2d_stats *stats) {
	if(!h) return fail(nullptr, QPDE_ERR_INVALID, "qpde_hjbqvi2d_solve: NULL handle");
	if(!g || !V) return fail(h, QPDE_ERR_INVALID, "qpde_hjbqvi2d_solve: NULL argument");
	if(g->abi_version != QPDE_ABI_VERSION) return fail(h, QPDE_ERR_INVALID, "problem.abi_version %d != %d", g->abi_version, QPDE_ABI_VERSION);
	if(g->model != QPDE_HJBQVI2D_OPTIMAL_CONSUMPTION) return fail(h, QPDE_ERR_UNSUPPORTED, "hjbqvi2d: unknown model %d", g->model);
	if(g->refine < 0 || g->refine > 10) return fail(h, QPDE_ERR_INVALID, "hjbqvi2d: refine out of range");
	if(!g->w_axis || g->n_w_axis < 3 || !g->b_axis || g->n_b_axis < 3 || !g->control_axis || g->n_control_axis < 1
			|| !g->impulse_axis || g->n_impulse_axis < 1) return fail(h, QPDE_ERR_INVALID, "hjbqvi2d: axes missing (the state axes need >= 3 ticks)");
	if(!g->params || g->n_params < 8) return fail(h, QPDE_ERR_INVALID, "hjbqvi2d: the optimal-consumption model takes 8 parameters");
	if(!(g->T > 0.) || g->timesteps < 1 || g->max_inner < 1 || g->max_sweeps < 1) return fail(h, QPDE_ERR_INVALID, "hjbqvi2d: T, timesteps, max_inner and max_sweeps must be positive");
	if(!(g->scaling_factor > 0.) || !(g->tolerance > 0.)) return fail(h, QPDE_ERR_INVALID, "hjbqvi2d: scaling_factor and tolerance must be > 0");
	const double *ax[4] = { g->w_axis, g->b_axis, g->control_axis, g->impulse_axis };
	const int len[4] = { g->n_w_axis, g->n_b_axis, g->n_control_axis, g->n_impulse_axis };
	for(int k = 0; k < 4; ++k) for(int i = 1; i < len[k]; ++i) if(!(ax[k][i] > ax[k][i - 1])) return fail(h, QPDE_ERR_INVALID, "hjbqvi2d: axis %d is not strictly increasing", k);
	if(!(g->params[5] >= 0. && g->params[5] < 1.)) return fail(h, QPDE_ERR_INVALID, "hjbqvi2d: the proportional cost lambda must lie in [0, 1)");
	if(!(g->params[4] > 0.)) return fail(h, QPDE_ERR_INVALID, "hjbqvi2d: gamma must be > 0");
	int size[4];
	std::vector<double> grid[4];
	for(int k = 0; k < 4; ++k) {
		size[k] = qpde_refined_size(len[k], g->refine);
		grid[k].resize((size_t)size[k]);
		qpde_refine_axis(ax[k], len[k], len[k] < 2 ? 0 : g->refine, grid[k].data());
	}
	if((long long)size[0] * size[1] > (1ll << 26)) return fail(h, QPDE_ERR_UNSUPPORTED, "hjbqvi2d: more than 2^26 nodes");
	QPDE_CUDA(h, cudaSetDevice(h->device));
	std::string error;
	const int rc = hjbqvi2d_run(g, grid[0].data(), size[0], grid[1].data(), size[1], grid[2].data(), size[2], grid[3].data(), size[3],
		V, Q, Z, mask, stats, h->stream, &h->launches, &error);
	if(rc != QPDE_OK) return fail(h, rc, "%s", error.c_str());
	if(W_out) memcpy(W_out, grid[0].data(), sizeof(double) * size[0]);
	if(B_out) memcpy(B_out, grid[1].data(), sizeof(double) * size[1]);
	return QPDE_OK;
}

int qpde_comm_unique_id(unsigned char id[QPDE_COMM_ID_BYTES]) {
	if(!id) return fail(nullptr, QPDE_ERR_INVALID, "qpde_comm_unique_id: NULL argument");
	std::string error;
	const int rc = comm_unique_id(id, &error);
	return rc == QPDE_OK ? rc : fail(nullptr, rc, "%s", error.c_str());
}

int qpde_comm_init(qpde_handle *h, const unsigned char id[QPDE_COMM_ID_BYTES], int rank, int world) {
	if(!h) return fail(nullptr, QPDE_ERR_INVALID, "qpde_comm_init: NULL handle");
	if(!id || world < 1 || rank < 0 || rank >= world) return fail(h, QPDE_ERR_INVALID, "qpde_comm_init: bad rank / world");
	if(h->comm) return fail(h, QPDE_ERR_INVALID, "qpde_comm_init: the handle already has a communicator");
	QPDE_CUDA(h, cudaSetDevice(h->device));
	std::string error;
	const int rc = comm_create(id, rank, world, &h->comm, &error);
	return rc == QPDE_OK ? rc : fail(h, rc, "%s", error.c_str());
}

int qpde_comm_destroy(qpde_handle *h) {
	if(!h) return QPDE_OK;
	cudaSetDevice(h->device);
	cudaStreamSynchronize(h->stream);
	comm_destroy(h->comm);
	h->comm = nullptr;
	return QPDE_OK;
}

static int gmwb_solve_impl(qpde_handle *h, const qpde_gmwb2d *g, double *V, double *W_out, double *A_out,
		qpde_gmwb2d_stats *stats, Comm *comm);

int qpde_gmwb2d_solve_sharded(qpde_handle *h, const qpde_gmwb2d *g, double *V, double *W_out, double *A_out,
		qpde_gmwb2d_stats *stats) {
	if(!h) return fail(nullptr, QPDE_ERR_INVALID, "qpde_gmwb2d_solve_sharded: NULL handle");
	if(!h->comm) return fail(h, QPDE_ERR_INVALID, "qpde_gmwb2d_solve_sharded: the handle has no communicator (qpde_comm_init)");
	return gmwb_solve_impl(h, g, V, W_out, A_out, stats, h->comm);
}

int qpde_gmwb2d_solve(qpde_handle *h, const qpde_gmwb2d *g, double *V, double *W_out, double *A_out,
		qpde_gmwb2d_stats *stats) {
	return gmwb_solve_impl(h, g, V, W_out, A_out, stats, nullptr);
}

static int gmwb_solve_impl(qpde_handle *h, const qpde_gmwb2d *g, double *V, double *W_out, double *A_out,
		qpde_gmwb2d_stats *stats, Comm *comm) {
	if(!h) return fail(nullptr, QPDE_ERR_INVALID, "qpde_gmwb2d_solve: NULL handle");
	if(!g || !V) return fail(h, QPDE_ERR_INVALID, "qpde_gmwb2d_solve: NULL argument");
	{ const int vrc = gmwb_validate(h, g); if(vrc != QPDE_OK) return vrc; }
	int nw = 0, na = 0;
	if(qpde_gmwb2d_nodes(g, &nw, &na) != QPDE_OK) return fail(h, QPDE_ERR_INVALID, "gmwb: bad grid");
	const int nz = qpde_refined_size(g->n_impulse_axis, g->refine);
	std::vector<double> W((size_t)nw), A((size_t)na), Z((size_t)nz);
	qpde_refine_axis(g->w_axis, g->n_w_axis, g->refine, W.data());
	qpde_refine_axis(g->a_axis, g->n_a_axis, g->refine, A.data());
	qpde_refine_axis(g->impulse_axis, g->n_impulse_axis, g->refine, Z.data());
	QPDE_CUDA(h, cudaSetDevice(h->device));
	std::string error;
	const int rc = gmwb_run(g, W.data(), nw, A.data(), na, Z.data(), nz, V, stats, h->stream, &h->launches, &error, comm);
	if(rc != QPDE_OK) return fail(h, rc, "%s", error.c_str());
	if(W_out) memcpy(W_out, W.data(), sizeof(double) * nw);
	if(A_out) memcpy(A_out, A.data(), sizeof(double) * na);
	return QPDE_OK;
}

} // extern "C" (reopened below)

struct qpde_gmwb2d_shard { qpde_handle *h; GmwbShard *impl; };

extern "C" {

int qpde_gmwb2d_shard_create(qpde_handle *h, const qpde_gmwb2d *g, int a_begin, int a_end, qpde_gmwb2d_shard **out) {
	if(!h) return fail(nullptr, QPDE_ERR_INVALID, "qpde_gmwb2d_shard_create: NULL handle");
	if(!g || !out) return fail(h, QPDE_ERR_INVALID, "qpde_gmwb2d_shard_create: NULL argument");
	*out = nullptr;
	{ const int vrc = gmwb_validate(h, g); if(vrc != QPDE_OK) return vrc; }
	int nw = 0, na = 0;
	if(qpde_gmwb2d_nodes(g, &nw, &na) != QPDE_OK) return fail(h, QPDE_ERR_INVALID, "gmwb: bad grid");
	if(a_begin < 0 || a_end > na || a_begin > a_end) return fail(h, QPDE_ERR_INVALID, "gmwb shard: lines [%d, %d) outside [0, %d)", a_begin, a_end, na);
	const int nz = qpde_refined_size(g->n_impulse_axis, g->refine);
	std::vector<double> W((size_t)nw), A((size_t)na), Z((size_t)nz);
	qpde_refine_axis(g->w_axis, g->n_w_axis, g->refine, W.data());
	qpde_refine_axis(g->a_axis, g->n_a_axis, g->refine, A.data());
	qpde_refine_axis(g->impulse_axis, g->n_impulse_axis, g->refine, Z.data());
	QPDE_CUDA(h, cudaSetDevice(h->device));
	std::string error;
	GmwbShard *impl = nullptr;
	const int rc = gmwb_shard_create(g, W.data(), nw, A.data(), na, Z.data(), nz, a_begin, a_end, h->stream,
		&h->launches, &impl, &error);
	if(rc != QPDE_OK) return fail(h, rc, "%s", error.c_str());
	*out = new qpde_gmwb2d_shard{h, impl};
	return QPDE_OK;
}

int qpde_gmwb2d_shard_destroy(qpde_gmwb2d_shard *s) {
	if(!s) return QPDE_OK;
	cudaSetDevice(s->h->device);
	cudaStreamSynchronize(s->h->stream);
	gmwb_shard_destroy(s->impl);
	delete s;
	return QPDE_OK;
}

#define QPDE_SHARD_CALL(call) \
	if(!s) return fail(nullptr, QPDE_ERR_INVALID, "gmwb shard: NULL shard"); \
	QPDE_CUDA(s->h, cudaSetDevice(s->h->device)); \
	call; \
	QPDE_CUDA(s->h, cudaGetLastError()); \
	return QPDE_OK

int qpde_gmwb2d_shard_exit(qpde_gmwb2d_shard *s, double *x_dev) { QPDE_SHARD_CALL(gmwb_shard_exit(s->impl, x_dev)); }
int qpde_gmwb2d_shard_policy(qpde_gmwb2d_shard *s, const double *x_dev, double h0) {
	QPDE_SHARD_CALL(gmwb_shard_policy(s->impl, x_dev, h0));
}
int qpde_gmwb2d_shard_sweep(qpde_gmwb2d_shard *s, double *x_dev, const double *v0_dev, int *status_dev) {
	QPDE_SHARD_CALL(QPDE_CUDA(s->h, gmwb_shard_sweep(s->impl, x_dev, v0_dev, status_dev)));
}
int qpde_gmwb2d_shard_relerr(qpde_gmwb2d_shard *s, const double *x_dev, const double *xprev_dev, int *notdone_dev) {
	QPDE_SHARD_CALL(gmwb_shard_relerr(s->impl, x_dev, xprev_dev, notdone_dev));
}
#undef QPDE_SHARD_CALL

int qpde_bs1d_plan_max_steps(const qpde_bs1d_plan *p) { return p ? p->b.max_steps : QPDE_ERR_INVALID; }

int qpde_bs1d_sweep(qpde_handle *h, const qpde_bs1d_batch *batch, qpde_bs1d_result *r) {
	if(!r) return fail(h, QPDE_ERR_INVALID, "qpde_bs1d_sweep: NULL result");
	int outputs = 0;
	if(r->V) outputs |= QPDE_OUT_V;
	if(r->S) outputs |= QPDE_OUT_S;
	if(r->value) outputs |= QPDE_OUT_VALUE;
	if(r->n_steps) outputs |= QPDE_OUT_STEPS;
	if(r->inner_total) outputs |= QPDE_OUT_INNER_TOTAL;
	if(r->inner) outputs |= QPDE_OUT_INNER;
	if(r->active) outputs |= QPDE_OUT_ACTIVE;
	if(r->status) outputs |= QPDE_OUT_STATUS;
	if(r->solves_computed) outputs |= QPDE_OUT_COMPUTED;
	qpde_bs1d_plan *p = nullptr;
	int rc = plan_create_impl(h, batch, outputs, &p);
	if(rc != QPDE_OK) return rc;
	const char *no_ranges = getenv("QPDE_SWEEP_NO_RANGES");      // tuning: one launch, copies after it
	if(p->kernel == QPDE_KERNEL_RESIDENT && !p->jump && p->b.C >= 16 * 2 * h->sm_count && !(no_ranges && no_ranges[0] == '1')) {
		QPDE_CUDA(h, cudaSetDevice(h->device));
		rc = run_and_fetch_ranges(p, r);
	} else {
		rc = qpde_bs1d_plan_run(p);
		if(rc == QPDE_OK) rc = qpde_bs1d_plan_fetch(p, r);
	}
	qpde_bs1d_plan_destroy(p);
	return rc;
}

} // extern "C"
