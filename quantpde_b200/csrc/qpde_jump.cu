// qpde_jump.cu — jump-diffusion sweep: one CTA per contract, the European
// backward sweep of BlackScholesJumpDiffusion1 under ReverseBDFOne / BDFTwo with
// the jump integral treated explicitly (d'Halluin, Forsyth & Vetzal's
// correlation-integral method), everything of one contract on chip.
//
// Reference call stack replaced:
//   BlackScholesJumpDiffusion::b      Modules/Operators/BlackScholes.hpp:389-459
//     V(exp(x0 + k dx)) by PiecewiseLinear   Core/Interpolant.hpp:153-181,331-374
//     Eigen::FFT fwd / inv (3 transforms of n_fft per step)
//     h(log S_i) by PiecewiseLinear on the uniform log-grid
//   BlackScholes::A with lambda, kappa       Modules/Operators/BlackScholes.hpp:136-226
//   BDFOne / BDFTwo ::A, ::b                 Core/BDF.hpp:32-117
//   SparseLUSolver                           Bindings/Eigen.hpp:143-165
//
// Per time step: (1) gather V onto the log-grid through a per-contract
// interpolation table, written bit-reversed into shared memory, two real points
// per complex element; (2) a hand-written radix-2 FFT pair of n_fft / 2 complex
// points in shared memory (forward decimation-in-time, one pass that unpacks the
// real-data spectrum, multiplies by the conjugate spectrum of the density weights
// — kept in shared memory — and packs for the way back, inverse
// decimation-in-frequency) — no cuFFT; (3) interpolate h back at log S_i and
// add h0 * lambda * h to the BDF right-hand side; (4) the partition solve with
// the cached factorisation (the matrix is constant in time).
//
// The circular, un-padded correlation of the reference is reproduced as is.
#include <math_constants.h>

#include "qpde_kernels.cuh"
#include "qpde_partition.cuh"

namespace qpde {

namespace {

constexpr int JUMP_ROW_ARRAYS = 5;  // S, lo, di, up, V

struct JumpSmem {
	PartitionSmem ps;
	struct Tail { double S, di, x, bd, invb, vprev, c, h; } tail;
};

// padded sizes of the complex work buffer and of the twiddle table (pad_buf, pad_tw below)
__host__ __device__ inline int pad_buf_size(int NF) { return NF + (NF >> 3) + 1; }
__host__ __device__ inline int pad_tw_size(int NF) { const int h = NF / 2; return h + (h >> 3) + (h >> 6) + (h >> 9) + 1; }

inline size_t jump_smem_bytes(int M, int P, int NF) {
	// row arrays [M][P] (+1 slot of V for the tail node); the transforms are of NF / 2 complex
	// points (real data, see below): work buffer, twiddles of the half-size transform,
	// exp(-2 pi i k / NF) for k <= NF / 4, spectrum of the density weights for k <= NF / 2
	const int NH = NF / 2;
	return sizeof(JumpSmem) + sizeof(double) * ((size_t)JUMP_ROW_ARRAYS * M * P + P + 2)
		+ sizeof(double2) * ((size_t)pad_buf_size(NH) + pad_tw_size(NH) + (NH / 2 + 1) + (NH + 1));
}

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
	return make_double2(fma(a.x, b.x, -(a.y * b.y)), fma(a.x, b.y, a.y * b.x));
}

// In-place radix-2 FFT of `buf` (NF complex points in shared memory) by the
// whole CTA.  forward = decimation in time: input bit-reversed, output natural,
// kernel exp(-2 pi i jk / NF).  inverse = decimation in frequency: input
// natural, output bit-reversed, kernel exp(+2 pi i jk / NF), unscaled.
// tw[k] = exp(-2 pi i k / NF), k < NF / 2.  Ends with a barrier.
//
// Three consecutive radix-2 stages are fused into one pass: a thread loads the
// 2^3 points that these stages connect (stride 2^s0), runs the butterflies in
// registers — the very operations of the one-stage-per-barrier form, in the same
// order, so the result is bit-identical to it — and stores them back: a third of
// the shared-memory traffic and of the barriers.  Both arrays are padded against
// bank conflicts (16-byte elements: a quarter-warp must hit 8 distinct 4-bank
// groups): the data by one element per 8 (the stride-1 pass), the twiddles by
// the base-8 digit sum of the index (8 threads read tw[j << t], j consecutive).
__device__ __forceinline__ int pad_buf(int i) { return i + (i >> 3); }
__device__ __forceinline__ int pad_tw(int k) { return k + (k >> 3) + (k >> 6) + (k >> 9); }

template <int U, int PT, bool INVERSE>
__device__ __forceinline__ void fft_pass(double2 *buf, const double2 *tw, int NF, int logNF, int s0, int tid) {
	constexpr int R = 1 << U;
	const int stride = 1 << s0;
	for(int g = tid; g < (NF >> U); g += PT) {
		const int low = g & (stride - 1);
		const int base = ((g >> s0) << (s0 + U)) | low;
		double2 v[R];
#pragma unroll
		for(int m = 0; m < R; ++m) v[m] = buf[pad_buf(base + m * stride)];
#pragma unroll
		for(int uu = 0; uu < U; ++uu) {
			const int u = INVERSE ? U - 1 - uu : uu;      // DIF runs the stages downwards
			const int half = 1 << u;
			const int shift = logNF - 1 - (s0 + u);
#pragma unroll
			for(int m = 0; m < R; ++m) {
				if(m & half) continue;
				// the pair (i, i + 2^s) of stage s = s0 + u with i = base + m stride; k = i mod 2^s
				const int k = low + (m & (half - 1)) * stride;
				const double2 w = tw[pad_tw(k << shift)];
				if(!INVERSE) {
					const double2 a = v[m];
					const double2 t = cmul(v[m + half], w);
					v[m] = make_double2(a.x + t.x, a.y + t.y);
					v[m + half] = make_double2(a.x - t.x, a.y - t.y);
				} else {
					const double2 a = v[m], c = v[m + half];
					const double2 d = make_double2(a.x - c.x, a.y - c.y);
					v[m] = make_double2(a.x + c.x, a.y + c.y);
					v[m + half] = cmul(d, make_double2(w.x, -w.y));    // conjugate twiddle
				}
			}
		}
#pragma unroll
		for(int m = 0; m < R; ++m) buf[pad_buf(base + m * stride)] = v[m];
	}
	__syncthreads();
}

template <int PT, bool INVERSE>
__device__ __forceinline__ void fft_pass_any(double2 *buf, const double2 *tw, int NF, int logNF, int s0, int U, int tid) {
	if(U == 3) fft_pass<3, PT, INVERSE>(buf, tw, NF, logNF, s0, tid);
	else if(U == 2) fft_pass<2, PT, INVERSE>(buf, tw, NF, logNF, s0, tid);
	else fft_pass<1, PT, INVERSE>(buf, tw, NF, logNF, s0, tid);
}

template <int PT>
__device__ __forceinline__ void fft_forward_dit(double2 *buf, const double2 *tw, int NF, int logNF, int tid) {
	for(int s0 = 0; s0 < logNF; s0 += 3) fft_pass_any<PT, false>(buf, tw, NF, logNF, s0, logNF - s0 < 3 ? logNF - s0 : 3, tid);
}

template <int PT>
__device__ __forceinline__ void fft_inverse_dif(double2 *buf, const double2 *tw, int NF, int logNF, int tid) {
	for(int s0 = 3 * ((logNF - 1) / 3); s0 >= 0; s0 -= 3) fft_pass_any<PT, true>(buf, tw, NF, logNF, s0, logNF - s0 < 3 ? logNF - s0 : 3, tid);
}

// The data are real: the NF real points x_n are transformed as NH = NF / 2 complex points
// z_m = x_2m + i x_2m+1.  With Z = DFT_NH(z), E / O the transforms of the even / odd samples and
// q_k = exp(-2 pi i k / NF):  E_k = (Z_k + conj Z_{NH-k}) / 2,  O_k = (Z_k - conj Z_{NH-k}) / (2 i),
// X_k = E_k + q_k O_k,  X_{NH-k} = conj(E_k - q_k O_k)  (k = 0 .. NH / 2; X_NH is the Nyquist term).
// The inverse of a Hermitian spectrum Y goes the same way back:  Z'_k = A_k + i B_k with
// A_k = Y_k + conj Y_{NH-k},  B_k = (Y_k - conj Y_{NH-k}) conj q_k,  Z'_{NH-k} = conj A_k + i conj B_k;
// the unscaled inverse DFT_NH of Z' is NF (y_2m + i y_2m+1).
// spectral_pass does, for the pair (k, NH - k) in place in `buf` (natural order):
//   STORE:   F_k, F_{NH-k} <- X_k, X_{NH-k}                       (set-up: spectrum of the weights)
//   else:    Y = X conj F, then Z' into buf                       (every step: the correlation)
template <int PT, bool STORE>
__device__ __forceinline__ void spectral_pass(double2 *buf, const double2 *tq, double2 *F, int NH, int tid) {
	for(int k = tid; k <= NH / 2; k += PT) {
		const int km = k == 0 ? 0 : NH - k;
		const double2 Zk = buf[pad_buf(k)], Zm = buf[pad_buf(km)];
		const double2 E = make_double2(.5 * (Zk.x + Zm.x), .5 * (Zk.y - Zm.y));
		const double2 O = make_double2(.5 * (Zk.y + Zm.y), -.5 * (Zk.x - Zm.x));
		const double2 q = tq[k];
		const double2 T = cmul(q, O);
		const double2 Xk = make_double2(E.x + T.x, E.y + T.y);
		const double2 Xm = make_double2(E.x - T.x, -(E.y - T.y));
		if(STORE) {
			F[k] = Xk; F[NH - k] = Xm;
		} else {
			const double2 Fk = F[k], Fm = F[NH - k];
			const double2 Yk = cmul(Xk, make_double2(Fk.x, -Fk.y));
			const double2 Ym = cmul(Xm, make_double2(Fm.x, -Fm.y));
			const double2 A = make_double2(Yk.x + Ym.x, Yk.y - Ym.y);
			const double2 D = make_double2(Yk.x - Ym.x, Yk.y + Ym.y);
			const double2 B = cmul(D, make_double2(q.x, -q.y));
			buf[pad_buf(k)] = make_double2(A.x - B.y, A.y + B.x);              // A + i B
			if(km != k && k != 0) buf[pad_buf(km)] = make_double2(A.x + B.y, -A.y + B.x);   // conj A + i conj B
		}
	}
	__syncthreads();
}

__device__ __forceinline__ int bitrev(int k, int logNF) { return (int)(__brev((unsigned)k) >> (32 - logNF)); }

// linearInterpolationData (Core/Interpolant.hpp:153-181) on ticks given by a functor
template <typename Tick>
__device__ __forceinline__ void interp_data(const Tick &x, int length, double ci, int &idx, double &w) {
	if(ci <= x(0)) { idx = 0; w = 1.; return; }
	if(ci >= x(length - 1)) { idx = length - 2; w = 0.; return; }
	int lo = 0, hi = length - 2, mid = 0;
	w = 0.;
	while(lo <= hi) {
		mid = (lo + hi) / 2;
		if(ci < x(mid)) hi = mid - 1;
		else if(ci >= x(mid + 1)) lo = mid + 1;
		else { w = (x(mid + 1) - ci) / (x(mid + 1) - x(mid)); break; }
	}
	idx = mid;
}

// PiecewiseLinear::interpolate for one dimension (:331-374): corners whose
// factor is <= epsilon are skipped
__device__ __forceinline__ double interp_apply(double w, double v0, double v1) {
	double sum = 0.;
	if(w > DBL_EPSILON) sum += w * v0;
	if(1. - w > DBL_EPSILON) sum += (1. - w) * v1;
	return sum;
}

} // namespace

// GT: forward-time sweeps (the common reverse sweep compiles the direction in, see ReplayT)
template <int M, int PT, bool BDF2, bool GT>
__global__ void __launch_bounds__(PT, 1)
k_jump_sweep(DeviceBatch b, JumpWork work, DeviceResult out) {
	constexpr int P = PT;
	constexpr int MP = M * P;
	extern __shared__ __align__(16) unsigned char smem_raw[];
	JumpSmem &sm = *reinterpret_cast<JumpSmem *>(smem_raw);
	double *const sm_S  = reinterpret_cast<double *>(smem_raw + sizeof(JumpSmem));
	double *const sm_lo = sm_S + MP;
	double *const sm_di = sm_lo + MP;
	double *const sm_up = sm_di + MP;
	double *const sm_V  = sm_up + MP;             // [MP + P + 1]: current solution in node order (at_node below)
	const int NF = b.n_fft, NH = NF / 2;          // NF real points = NH complex points (spectral_pass)
	const int logNH = 31 - __clz(NH);
	double2 *const sm_buf = reinterpret_cast<double2 *>(sm_V + MP + P + 2);   // [NH], padded (pad_buf)
	double2 *const sm_tw = sm_buf + pad_buf_size(NH);                     // [NH / 2], padded (pad_tw): exp(-2 pi i k / NH)
	double2 *const sm_tq = sm_tw + pad_tw_size(NH);                       // [NH / 2 + 1]: exp(-2 pi i k / NF)
	double2 *const sm_F = sm_tq + (NH / 2 + 1);                           // [NH + 1]: spectrum of the density weights

	const int cidx = blockIdx.x;
	const int tid = threadIdx.x;
	const int lane = tid & 31, warp = tid >> 5;
	const int N = b.N;
	const bool has_tail = N == P * M + 1;
	const int n_main = has_tail ? N - 1 : N;
	const int i0 = tid * M;
	const bool tail_owner = has_tail && tid == P - 1;

	const double K = b.strike[cidx];
	const double lambda = b.jump_lambda[cidx], kappa = b.jump_kappa[cidx];
	const double x0 = b.jump_x0[cidx], dxg = b.jump_dx[cidx];
	int *const g_idx = work.g_idx + (size_t)cidx * NF;
	double *const g_w = work.g_w + (size_t)cidx * NF;
	int *const n_idx = work.n_idx + (size_t)cidx * N;
	double *const n_w = work.n_w + (size_t)cidx * N;

	// node i of V / S in shared memory
	// The log-grid gather reads V at neighbouring nodes from neighbouring lanes: V is kept in node order, so that
	// they fall into neighbouring banks (in the [row][thread] order of the other row arrays, nodes i and i + 1 are
	// P doubles apart — the same bank: 38 % of the kernel's shared-memory wavefronts were conflicts).  One pad per
	// thread when M is even keeps the threads' own stores (M rows each) conflict-free too.
	auto at_node = [&] (int i) -> int { return i >= n_main ? MP + P : (M % 2 == 0 ? i + i / M : i); };

	// ------------------------------------------------------------------
	// Setup
	// ------------------------------------------------------------------
	{
		const double *axis = b.axis + (size_t)cidx * b.axis_stride;
		for(int j = 0; j < M; ++j) {
			const int i = i0 + j;
			sm_S[j * P + tid] = i < n_main ? refined_node(axis, b.refine, i) : 0.;
		}
		if(tid == 0) sm.tail.S = has_tail ? refined_node(axis, b.refine, N - 1) : 0.;
		for(int k = tid; k <= NH / 2; k += P) {
			double sn, cs;
			sincospi(-2. * (double)k / (double)NF, &sn, &cs);
			sm_tq[k] = make_double2(cs, sn);
			if(k < NH / 2) {
				sincospi(-2. * (double)k / (double)NH, &sn, &cs);
				sm_tw[pad_tw(k)] = make_double2(cs, sn);
			}
		}
	}
	__syncthreads();
	auto S_at = [&] (int i) -> double { return i >= n_main ? sm.tail.S : sm_S[(i % M) * P + (i / M)]; };
	{
		const double r = b.r[cidx], q = b.q[cidx], sig = b.sigma[cidx];
		for(int j = 0; j <= M; ++j) {
			const bool tail = j == M;
			if(tail && (tid != 0 || !has_tail)) break;
			const int i = tail ? N - 1 : i0 + j;
			Row row; row.lo = 0.; row.di = 0.; row.up = 0.;
			if(tail || i < n_main) {
				const double Si = S_at(i);
				const double Sm = i > 0 ? S_at(i - 1) : 0.;
				const double Sp = i < N - 1 ? S_at(i + 1) : 0.;
				const double v_i = b.vol_model == QPDE_VOL_NODES
					? b.sigma_nodes[(size_t)cidx * b.sigma_nodes_stride + i]
					: vol_value(b.vol_model, sig, Si);
				row = bs_row(i, N, Sm, Si, Sp, r, v_i, q, lambda, kappa);
			}
			if(tail) sm.tail.di = row.di;
			else { sm_lo[j * P + tid] = row.lo; sm_di[j * P + tid] = row.di; sm_up[j * P + tid] = row.up; }
		}
		// interpolation tables
		// (a) V at exp(x0 + k dx): PiecewiseLinear over the S grid (BlackScholes.hpp:405-414)
		for(int k = tid; k < NF; k += P) {
			const double c = exp(x0 + k * dxg);
			int idx; double w;
			interp_data(S_at, N, c, idx, w);
			g_idx[k] = at_node(idx) | (at_node(idx + 1) << 16);      // both shared-memory slots (< 2^16), no division per step
			g_w[k] = w;
		}
		// (b) h at log S_i: PiecewiseLinear over Axis::uniform(x0, xf, NF) (:293-296, :441-452)
		const double xf = log(S_at(N - 2));
		const double du = (xf - x0) / (NF - 1);        // Axis::uniform, Core/Axis.hpp:219-226
		auto tick = [&] (int k) -> double { return x0 + du * k; };
		for(int i = tid; i < N; i += P) {
			int idx = 0; double w = 1.;
			if(i > 0 && i < N - 1) interp_data(tick, NF, log(S_at(i)), idx, w);
			// h_k is the real (k even) or imaginary (k odd) part of element bitrev(k / 2) of sm_buf:
			// slots in sm_buf viewed as doubles, for idx and idx + 1
			n_idx[i] = (2 * pad_buf(bitrev(idx >> 1, logNH)) + (idx & 1))
				| ((2 * pad_buf(bitrev((idx + 1) >> 1, logNH)) + ((idx + 1) & 1)) << 16);
			n_w[i] = i > 0 && i < N - 1 ? w : 0.;                     // rows 0 and N - 1 have b = 0
		}
		// (c) spectrum of the density weights (computeDensityFFT, :326-328)
		const double *wts = b.jump_weights + (size_t)cidx * b.jump_weights_stride;
		for(int m = tid; m < NH; m += P) sm_buf[pad_buf(bitrev(m, logNH))] = make_double2(wts[2 * m], wts[2 * m + 1]);
	}
	__syncthreads();
	fft_forward_dit<PT>(sm_buf, sm_tw, NH, logNH, tid);
	spectral_pass<PT, true>(sm_buf, sm_tq, sm_F, NH, tid);     // ends with a barrier

	// ---- per-thread state ------------------------------------------------
	double x[M], vprev[BDF2 ? M : 1];
	double x_next;
	PartitionSolver<M, PT> ps; ps.init();
	auto initial_value = [&] (int i) -> double {
		return b.payoff_kind == QPDE_PAYOFF_ARRAY
			? b.payoff[(size_t)cidx * b.payoff_stride + i]
			: payoff_value(b.payoff_kind, S_at(i), K);
	};
#pragma unroll
	for(int j = 0; j < M; ++j) {
		x[j] = i0 + j < n_main ? initial_value(i0 + j) : 0.;
		if(BDF2) vprev[j] = 0.;
	}
	if(tail_owner) { sm.tail.x = initial_value(N - 1); sm.tail.vprev = 0.; sm.tail.c = 0.; sm.tail.bd = 1.; sm.tail.invb = 1.; sm.tail.h = 0.; }
	x_next = 0.;
	__syncthreads();

	ReplayT<GT> rp; rp.start(b.T[cidx], b.dt[cidx], 0, b.forward);
	NodeState ns; ns.start(b.scheme);
	double time1 = rp.initial(), time0 = rp.initial();
	int cur_kind = -1; double cur_h = 0.;
	int n_steps = 0;
	bool more = true;

	while(more) {
		qpde_step st;
		more = rp.next(st);
		const double t = st.t;
		const int kind = ns.kind();
		const double h0 = rp.lapse(time1, t);
		Gamma g; g.kind = kind; g.h = h0;
		double c1 = 0., c0 = 0., den = 1.;
		if(BDF2 && kind == STEP_BDF2) {
			const double h1 = rp.lapse(time1, t), hh0 = rp.lapse(time0, t);
			g.h = div_step(hh0 * h1, hh0 + h1);
			c1 = div_step(hh0, h1); c0 = div_step(h1, hh0); den = hh0 - h1;
		}
		const double rden = div_step(1., den);
		// the matrix I + g.h * A is constant in time: factorise when the step changes
		if(kind != cur_kind || fabs(g.h - cur_h) > 1e-11 * cur_h) {
			cur_kind = kind; cur_h = g.h;
			double aa[M], bb[M], cc[M];
#pragma unroll
			for(int j = 0; j < M; ++j) {
				const int at = j * P + tid;
				aa[j] = g.off(sm_lo[at]); bb[j] = 1. + g.off(sm_di[at]); cc[j] = g.off(sm_up[at]);
				if(j == M - 1 && tail_owner) { sm.tail.c = cc[j]; cc[j] = 0.; }
			}
			if(tail_owner) { sm.tail.bd = 1. + g.off(sm.tail.di); sm.tail.invb = rcp(sm.tail.bd); }
			ps.factorize(aa, bb, cc, sm.ps, lane, warp);
		}

		// ---- (1) V onto the log-grid, bit-reversed ------------------------
#pragma unroll
		for(int j = 0; j < M; ++j) sm_V[M % 2 == 0 ? i0 + j + tid : i0 + j] = x[j];
		if(tail_owner) sm_V[MP + P] = sm.tail.x;
		__syncthreads();
		// two real points per complex element; the table loads of two elements are issued together
		for(int m0 = tid; m0 < NH; m0 += 2 * P) {
			int2 slots[2]; double2 wk[2];
#pragma unroll
			for(int u = 0; u < 2; ++u) {
				const int m = m0 + u * P < NH ? m0 + u * P : m0;
				slots[u] = reinterpret_cast<const int2 *>(g_idx)[m];
				wk[u] = reinterpret_cast<const double2 *>(g_w)[m];
			}
#pragma unroll
			for(int u = 0; u < 2; ++u) {
				const int m = m0 + u * P;
				const double v0 = interp_apply(wk[u].x, sm_V[slots[u].x & 0xffff], sm_V[slots[u].x >> 16]);
				const double v1 = interp_apply(wk[u].y, sm_V[slots[u].y & 0xffff], sm_V[slots[u].y >> 16]);
				if(m < NH) sm_buf[pad_buf(bitrev(m, logNH))] = make_double2(v0, v1);
			}
		}
		__syncthreads();
		// ---- (2) correlation: inv( fwd(Vbar) * conj(fwd(weights)) ), real data on half-size transforms
		fft_forward_dit<PT>(sm_buf, sm_tw, NH, logNH, tid);
		spectral_pass<PT, false>(sm_buf, sm_tq, sm_F, NH, tid);    // ends with a barrier
		// the read-out table of step (3): loaded here so that the inverse transform hides the latency
		int nslot[M]; double nw[M];
#pragma unroll
		for(int j = 0; j < M; ++j) {
			const int i = i0 + j < N ? i0 + j : 0;
			nslot[j] = n_idx[i]; nw[j] = n_w[i];
		}
		fft_inverse_dif<PT>(sm_buf, sm_tw, NH, logNH, tid);
		// h_k / NF: real / imaginary part of element bitrev(k / 2)
		const double *const sm_h = reinterpret_cast<const double *>(sm_buf);
		const double inv_nf = 1. / (double)NF;
		// ---- (3) right-hand side: BDF b() with the explicit jump term ------
		double r[M];
#pragma unroll
		for(int j = 0; j < M; ++j) {
			const int i = i0 + j;
			// no predicate across the loads: rows without a jump term have weight and slots 0
			const int slots = nslot[j];
			const double hA = sm_h[slots & 0xffff] * inv_nf;
			const double hB = sm_h[slots >> 16] * inv_nf;
			const double bj = i > 0 && i < N - 1 && i < n_main ? lambda * interp_apply(nw[j], hA, hB) : 0.;
			if(kind == STEP_IMPLICIT) r[j] = x[j] + h0 * bj;                                   // BDF.hpp:40-46
			else r[j] = (div_by(c1 * x[j] - c0 * vprev[BDF2 ? j : 0], den, rden) + bj) * g.h;             // BDF.hpp:94-110
		}
		double xt_new = 0.;
		if(tail_owner) {
			// the last row has b = 0 (BlackScholes.hpp:455-456)
			const double lbt = kind == STEP_IMPLICIT ? sm.tail.x + h0 * 0.
				: (div_by(c1 * sm.tail.x - c0 * sm.tail.vprev, den, rden) + 0.) * g.h;
			xt_new = lbt * sm.tail.invb;
			r[M - 1] = fma(-sm.tail.c, xt_new, r[M - 1]);
		}
		if(BDF2) {
#pragma unroll
			for(int j = 0; j < M; ++j) vprev[j] = x[j];
			if(tail_owner) sm.tail.vprev = sm.tail.x;
		}
		// ---- (4) solve ------------------------------------------------------
		ps.solve(r, x, x_next, sm.ps, lane, warp);
		if(tail_owner) sm.tail.x = xt_new;
		__syncthreads();     // sm_buf / sm_V are rewritten by the next step

		++n_steps;
		time0 = time1; time1 = t;
		ns.end_step();
		if(st.event_after) ns.after_event();
	}

	// ------------------------------------------------------------------
	// Outputs
	// ------------------------------------------------------------------
#pragma unroll
	for(int j = 0; j < M; ++j) {
		const int i = i0 + j;
		if(i < n_main) {
			if(out.V) out.V[(size_t)cidx * out.V_stride + i] = x[j];
			if(out.S) out.S[(size_t)cidx * out.S_stride + i] = sm_S[j * P + tid];
			if(out.active) out.active[(size_t)cidx * out.active_stride + i] = 0;
		}
	}
	if(tail_owner) {
		if(out.V) out.V[(size_t)cidx * out.V_stride + N - 1] = sm.tail.x;
		if(out.S) out.S[(size_t)cidx * out.S_stride + N - 1] = sm.tail.S;
		if(out.active) out.active[(size_t)cidx * out.active_stride + N - 1] = 0;
	}
	if(tid == 0) {
		if(out.n_steps) out.n_steps[cidx] = n_steps;
		if(out.inner_total) out.inner_total[cidx] = n_steps;
		if(out.computed) out.computed[cidx] = n_steps;
		if(out.status) out.status[cidx] = 0;
		if(out.inner) for(int s = 0; s < n_steps && s < b.max_steps; ++s) out.inner[(size_t)cidx * out.inner_stride + s] = 1;
	}
	if(out.value) {
		const double s0 = b.spot ? b.spot[cidx] : K;
		const double Sfirst = S_at(0), Slast = S_at(N - 1);
#pragma unroll
		for(int j = 0; j < M; ++j) {
			const int i = i0 + j;
			if(i >= N - 1) continue;
			const double Si = sm_S[j * P + tid];
			const double Sp = S_at(i + 1);
			const double vi = x[j];
			const double vp = i + 1 >= n_main ? sm.tail.x : (j == M - 1 ? x_next : x[j + 1 < M ? j + 1 : j]);
			bool mine = false; double wgt = 0.;
			if(s0 <= Sfirst) { mine = i == 0; wgt = 1.; }
			else if(s0 >= Slast) { mine = i == N - 2; wgt = 0.; }
			else if(!(s0 < Si) && !(s0 >= Sp)) { mine = true; wgt = (Sp - s0) / (Sp - Si); }
			if(mine) out.value[cidx] = interp_apply(wgt, vi, vp);
		}
	}
}

static int jump_geometry(int N, int *M, int *PT) {
	if(N <= 8 * 64 + 1)  { *M = 8; *PT = 64;  return 1; }
	if(N <= 8 * 256 + 1) { *M = 8; *PT = 256; return 1; }
	if(N <= 9 * 256 + 1) { *M = 9; *PT = 256; return 1; }
	return 0;
}

bool jump_supported(const DeviceBatch &b) {
	int M, PT;
	if(!b.jump_lambda || b.N < 10 || !jump_geometry(b.N, &M, &PT)) return false;
	if(b.american || b.n_controls > 0 || b.n_exercise > 0) return false;
	if(b.scheme != QPDE_SCHEME_BDF1 && b.scheme != QPDE_SCHEME_BDF2) return false;
	if(b.n_fft < 2 || (b.n_fft & (b.n_fft - 1)) != 0 || b.n_fft < b.N || b.n_fft > 4096) return false;
	return jump_smem_bytes(M, PT, b.n_fft) <= 227 * 1024;
}

template <int M, int PT, bool BDF2>
static int launch_jump_variant(const DeviceBatch &b, const JumpWork &w, const DeviceResult &out, cudaStream_t stream) {
	auto kern = b.forward ? k_jump_sweep<M, PT, BDF2, true> : k_jump_sweep<M, PT, BDF2, false>;
	const size_t bytes = jump_smem_bytes(M, PT, b.n_fft);
	if(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess) {
		return QPDE_ERR_CUDA;
	}
	kern<<<b.C, PT, bytes, stream>>>(b, w, out);
	return QPDE_OK;
}

int launch_jump(const DeviceBatch &b, const JumpWork &w, const DeviceResult &out,
		cudaStream_t stream, long long *launches) {
	int M = 0, PT = 0;
	if(!jump_supported(b) || !jump_geometry(b.N, &M, &PT)) return QPDE_ERR_UNSUPPORTED;
	const bool bdf2 = b.scheme == QPDE_SCHEME_BDF2;
	int rc;
	if(M == 8 && PT == 64)       rc = bdf2 ? launch_jump_variant<8, 64, true>(b, w, out, stream) : launch_jump_variant<8, 64, false>(b, w, out, stream);
	else if(M == 8 && PT == 256) rc = bdf2 ? launch_jump_variant<8, 256, true>(b, w, out, stream) : launch_jump_variant<8, 256, false>(b, w, out, stream);
	else                         rc = bdf2 ? launch_jump_variant<9, 256, true>(b, w, out, stream) : launch_jump_variant<9, 256, false>(b, w, out, stream);
	if(rc == QPDE_OK) *launches += 1;
	return rc;
}

} // namespace qpde
