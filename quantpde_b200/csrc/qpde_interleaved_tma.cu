// qpde_interleaved_tma.cu — the INTERLEAVED family with its HBM streams staged by TMA.
//
// Same algorithm, same arithmetic and same [C / 32][N][32] workspace as qpde_interleaved.cu (one
// thread per contract, Thomas elimination, the whole sweep in one launch), i.e. the same
// replacement of the reference call stack
//   TimeIteration<false>::iterateUntilDone     Core/IterativeMethod.hpp:1163-1204,1259-1317
//   ToleranceIteration + relativeError         Core/IterativeMethod.hpp:1125-1137,1211-1254
//   PenaltyMethodDifference::onIterationStart  Core/PenaltyMethod.hpp:37-69,96-119
//   Rannacher / CrankNicolson / BDFOne / BDFTwo A(), b()
//   SparseLUSolver::initialize/solve           Bindings/Eigen.hpp:143-165
// What changes is how the rows reach the serial chain.  A thread that loads its own rows
// pays the HBM latency once per block of rows and needs registers for everything in
// flight; here a warp of 32 contracts owns a ring of shared-memory stages that the TMA
// unit fills with tiles of 32 contracts x 4 node rows per array (1 KB, contiguous; one
// cp.async.bulk.tensor per group of adjacent arrays and tile, issued by lane 0, completion
// on an mbarrier), so that `stages` tiles are in flight per warp while the chain runs, at
// no register cost.  All control flow that surrounds a tile is warp-uniform (votes); what differs
// per contract (step kind, number of penalty iterations, number of steps) is a
// predicate on the stores.
//
// Passes per time step: the first solve forms the right-hand side on the way (reads lo, di,
// up, x [+ obstacle, xprev], writes lb, cp, dp [+ xprev]), every further penalty iteration
// has a forward elimination (reads lo, di, up, lb [+ obstacle, x], writes cp, dp), and every
// solve a back substitution (reads cp, dp [+ obstacle, x], writes x).  A penalty iteration
// whose active set equals the previous one's is counted, not computed, and a further penalty
// iteration starts its forward elimination at the first row whose active-set bit changed (minimum
// over the warp's contracts that iterate): cp, dp of the rows before it are those of the previous
// elimination bit for bit (same rows, same lb, same penalty terms).
// Tiles whose obstacle is zero for every contract of the warp (a put above its strike, a call below it) are
// loaded without the obstacle array (zero_tile below): 8 B per row and pass less on about half of the rows.
// (Tried and dropped: re-generating a put / call obstacle per row from the coarse axis instead of
// streaming it — 8 B per row and pass less, but no faster at 65,536 x 2049 and 1.5x slower at
// 16,384 contracts, where a warp has a scheduler to itself and every added instruction is latency.)
// Results written with ordinary stores are read by later TMA loads: every pass starts
// with a device-scope fence + fence.proxy.async.
#include <cuda.h>
#include <math_constants.h>

#include <cstdlib>

#include "qpde_kernels.cuh"
#include "qpde_partition.cuh"   // rcp()

namespace qpde {

namespace {

constexpr int TR = QPDE_TMA_TILE_ROWS;
constexpr int SLOT_DOUBLES = TR * 32;            // one array's tile: TR rows x 32 contracts
constexpr int MAX_SLOTS = 6;
constexpr int STAGE_DOUBLES = MAX_SLOTS * SLOT_DOUBLES;
constexpr int STAGE_BYTES = STAGE_DOUBLES * 8;   // 6 KB
constexpr unsigned FULL = 0xffffffffu;

// The workspace block is a stack of arrays; the tensor maps see it as [array][rows][32] with
// boxes of 32 contracts x TR rows x K arrays (K = 1, 2, 3): arrays that a pass reads together are
// adjacent in the stack — (lo, di, up), (obstacle, x), (cp, dp) — and arrive with ONE
// cp.async.bulk.tensor instead of one per array (the issue by lane 0 costs every lane's slots).
// These are the positions in the stack (-1 = not allocated).
struct TmaArrays { int coef, lb, obx, x, cpdp, xp; };
struct TmaMaps { CUtensorMap k[3]; };

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
	uint32_t done;
	do {
		asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
			: "=r"(done) : "r"(bar), "r"(parity) : "memory");
	} while(!done);
}

// one tile: box {32 contracts, TR rows, K arrays} at (row, array) of the workspace block
__device__ __forceinline__ void tma_tile(uint32_t dst, const CUtensorMap *map, int row, int array, uint32_t bar) {
	asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
		:: "r"(dst), "l"((unsigned long long)map), "r"(0), "r"(row), "r"(array), "r"(bar) : "memory");
}

// A pass reads up to three groups of adjacent arrays: group g = K arrays from stack position
// `array` into slots slot .. slot + K - 1 of the stage (K = 0: unused).
struct Boxes { int k[3], array[3], slot[3]; };

// The ring of one warp.  Every pass issues and consumes the same number of tiles, in the
// same order, so one stage cursor per side and one parity bit per stage are all the state.
struct Ring {
	uint32_t data, bars;   // shared-memory addresses of stage 0 / barrier 0
	int ns;                // stages
	int is, cs;            // stage of the next issue / of the next wait
	uint32_t parity;       // bit s: parity the next wait on stage s looks for
};

// Issues the tile at `row` (relative to this warp's block of contracts) for every group of `bx`.
__device__ __forceinline__ void ring_issue(Ring &r, const TmaMaps &maps, int blk0, const Boxes &bx, int row, int lane) {
	if(lane == 0) {
		const uint32_t bar = r.bars + 8u * r.is;
		const uint32_t dst = r.data + (uint32_t)STAGE_BYTES * r.is;
		mbar_expect_tx(bar, (uint32_t)((bx.k[0] + bx.k[1] + bx.k[2]) * SLOT_DOUBLES * 8));
#pragma unroll
		for(int g = 0; g < 3; ++g) {
			if(bx.k[g] > 0) tma_tile(dst + (uint32_t)(bx.slot[g] * SLOT_DOUBLES * 8), &maps.k[bx.k[g] - 1], blk0 + row, bx.array[g], bar);
		}
	}
	r.is = r.is + 1 == r.ns ? 0 : r.is + 1;
}

// Waits for the next tile; returns its stage.
__device__ __forceinline__ int ring_wait(Ring &r) {
	const int s = r.cs;
	mbar_wait(r.bars + 8u * s, (r.parity >> s) & 1u);
	r.parity ^= 1u << s;
	r.cs = s + 1 == r.ns ? 0 : s + 1;
	return s;
}

// Ordinary stores of this warp (and of earlier kernels) become visible to the TMA loads
// that follow.
__device__ __forceinline__ void pass_begin() {
	__threadfence();
	asm volatile("fence.proxy.async;" ::: "memory");
	__syncwarp();
}

// AMERICAN (the penalty iteration) is a compile-time variant: the box depths of every pass are then
// constants, and the linear sweep carries none of the obstacle code.
template <bool AMERICAN>
__global__ void __launch_bounds__(32, 8)
k_interleaved_sweep_tma(const __grid_constant__ TmaMaps maps, DeviceBatch b, InterleavedWork w, DeviceResult out,
		TmaArrays ta, int stages) {
	extern __shared__ __align__(128) unsigned char smem_raw[];
	const int lane = threadIdx.x;
	const int c = blockIdx.x * 32 + lane;
	const int blk0 = blockIdx.x * b.N;        // first tensor-map row of this warp's block of contracts inside an array
	const bool valid = c < b.C;
	const int cc = valid ? c : b.C - 1;       // lanes past the batch read a real contract's parameters and store nothing
	const int N = b.N;
	constexpr size_t ld = QPDE_WARP_ROW;      // row stride inside the block ([C / 32][N][32] layout)
	const size_t base = interleaved_base(c, N);
	const double tol = b.tolerance, scale = b.scale;
	const double pen = 1. / b.penalty_tolerance; // PenaltyMethod.hpp:85
	double *x = w.x + base, *xp = w.xprev + base, *lb = w.lb + base;
	double *cp = w.cp + base, *dp = w.dp + base;
	const double *ob = w.obst ? w.obst + base : nullptr;
	const double *evob = w.evobst ? w.evobst + base : nullptr;
	const double *sm = reinterpret_cast<const double *>(smem_raw) + lane;

	Ring ring;
	ring.data = smem_u32(smem_raw);
	ring.bars = ring.data + (uint32_t)(STAGE_BYTES * stages);
	ring.ns = stages; ring.is = 0; ring.cs = 0; ring.parity = 0u;
	if(lane == 0) {
		for(int s = 0; s < stages; ++s) mbar_init(ring.bars + 8u * s, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncwarp();

	constexpr bool american = AMERICAN;
	constexpr bool iterate = american;             // a ToleranceIteration wraps every step
	// Rows [0, z_lo) and [z_hi, N) carry a zero obstacle in every contract of this warp: their tiles are loaded
	// without the obstacle array.  One pass over the obstacle per sweep finds the two bounds.
	int z_lo = 0, z_hi = N;
	if(american) {
		int first_nz = N, last_nz = -1;
		for(int i = 0; i < N; ++i) {
			if(ob[(size_t)i * ld] != 0.) { if(first_nz == N) first_nz = i; last_nz = i; }
		}
		z_lo = __reduce_min_sync(FULL, valid ? first_nz : N);
		z_hi = __reduce_max_sync(FULL, valid ? last_nz : -1) + 1;
	}
	auto zero_tile = [&](int row0) -> bool { return american && (row0 + TR <= z_lo || row0 >= z_hi) && row0 >= 0; };
	const bool copy_xp = b.scheme == QPDE_SCHEME_BDF2 || b.variable_target != nullptr;   // iterand(1) <- iterand(0)
	const int nb = (N + TR - 1) / TR;            // tiles per pass
	const int lead = nb < stages ? nb : stages;

	Replay rp; rp.start(b.T[cc], b.dt[cc], b.n_exercise, b.forward, b.event_times);
	NodeState ns; ns.start(b.scheme);
	double time1 = rp.initial(), time0 = rp.initial(); // times of iterand(0), iterand(1)
	Gamma gA; gA.kind = STEP_IMPLICIT; gA.h = 0.; // what the factorised A was built with
	int n_steps = 0, inner_total = 0, computed = 0, status = 0;
	bool more = valid;
	// ReverseVariableStepper (Core/Stepper.hpp:60-126): every contract has its own step sizes
	// and its own number of steps; V^n is kept in xprev for the relative change over a step
	const bool variable = b.variable_target != nullptr;
	const double vtarget = variable ? b.variable_target[cc] : 0.;
	double vdt = b.dt[cc];

	while(__any_sync(FULL, more)) {
		if(variable && more) {
			if(n_steps >= b.max_steps) { status = 3; more = false; }     // output capacity
			rp.dt_const = vdt;                                            // VariableStepper::timestep()
		}
		const bool act = more;                   // this contract takes a step in this trip
		qpde_step st; st.t = 0.; st.dt = 0.; st.same_dt = 0; st.event_after = 0; st.user_events = 0; st.reserved = 0;
		int kind = STEP_IMPLICIT;
		double t = 0., h0 = 0., c1 = 0., c0 = 0., den = 1.;
		Gamma gb; gb.kind = STEP_IMPLICIT; gb.h = 0.;
		if(act) {
			more = rp.next(st);
			t = st.t;
			kind = ns.kind();
			h0 = rp.lapse(time1, t);
			gb.kind = kind; gb.h = h0;
			if(kind == STEP_BDF2) {
				const double h1 = rp.lapse(time1, t), hh0 = rp.lapse(time0, t);     // BDF.hpp:68-73
				gb.h = div_step(hh0 * h1, hh0 + h1);
				c1 = div_step(hh0, h1); c0 = div_step(h1, hh0); den = hh0 - h1;
			}
			// IterativeMethod.hpp:906: re-assemble A unless the root says it is the same
			// (a penalty node in the tree keeps LinearSystem's default: never)
			if(iterate || !ns.initialized || !ns.a_same(st.same_dt)) gA = gb;
		}
		const double rden = div_step(1., den);

		// One row of the elimination of (lA + P) x = lb + P obstacle (the serial chain)
		const double hsA = gA.factor(), hsb = gb.factor();   // Gamma::off(l) == l * factor(), without the switch
		double cprev = 0., dprev = 0.;
		auto eliminate = [&](int i, double vlo, double vdi, double vup, double rhs, double vob, double vx, bool store) {
			double a_i = vlo * hsA;
			double b_i = 1. + vdi * hsA;
			double c_i = vup * hsA;
			if(american) {
				// PenaltyMethod.hpp:45,61-68: predicate on the previous iterand
				// (scale * (x - payoff)) < 0  <=>  x < payoff
				if(vx < vob) { b_i = b_i + pen * 1.; rhs = rhs + pen * vob; }
			}
			const double denom = b_i - a_i * cprev;
			const double inv = rcp(denom);        // < 1 ulp; no division sequence on the serial chain
			cprev = c_i * inv;
			dprev = (rhs - a_i * dprev) * inv;
			if(store) { cp[(size_t)i * ld] = cprev; dp[(size_t)i * ld] = dprev; }
		};

		int its = 0;
		int fstart = 0;                          // first row of the next forward elimination
		bool again = act;
		bool first = true;
		do {
			if(first) {
				// ---- first solve of the step: lb = root.b(t) of the discretisation node is formed
				// on the way (and kept for the later penalty iterations), iterand(1) <- iterand(0)
				// for BDF2, and the forward elimination runs in the same pass.  Row i of lb needs
				// x[i + 1]: the chain runs one row behind the tiles.
				first = false;
				const bool any_b2 = __any_sync(FULL, act && kind == STEP_BDF2);
				// slots: lo di up | obstacle x | xprev
				const Boxes bx = { { 3, american ? 2 : 1, any_b2 ? 1 : 0 }, { ta.coef, american ? ta.obx : ta.x, ta.xp }, { 0, american ? 3 : 4, 5 } };
				const Boxes bz = { { 3, 1, any_b2 ? 1 : 0 }, { ta.coef, ta.x, ta.xp }, { 0, 4, 5 } };      // zero-obstacle tiles
				pass_begin();
				for(int j = 0; j < lead; ++j) { const int row_ = j * TR; if(zero_tile(row_)) ring_issue(ring, maps, blk0, bz, row_, lane); else ring_issue(ring, maps, blk0, bx, row_, lane); }
				double xm = 0., xc = 0., pxp = 0., plo = 0., pdi = 0., pup = 0., pob = 0.;
				cprev = 0.; dprev = 0.;
				auto finish_row = [&](int i, double vp) {
					double val;
					if(kind == STEP_IMPLICIT) {
						val = xc + 0. * h0;
					} else if(kind == STEP_BDF2) {
						val = (div_by(c1 * xc - c0 * pxp, den, rden) + 0.) * gb.h;
					} else {
						// (I - A h/2) v0, row-major SpMV from zero in column order
						double acc = 0.;
						if(i > 0) acc += (0. - plo * hsb) * xm;
						acc += (1. - pdi * hsb) * xc;
						if(i < N - 1) acc += (0. - pup * hsb) * vp;
						val = acc + 0.;
					}
					if(act) {
						lb[(size_t)i * ld] = val;
						if(copy_xp) xp[(size_t)i * ld] = xc;
					}
					eliminate(i, plo, pdi, pup, val, pob, xc, again);
				};
				for(int j = 0; j < nb; ++j) {
					const int s = ring_wait(ring);
					const double *tile = sm + s * STAGE_DOUBLES;
					double vlo[TR], vdi[TR], vup[TR], vx[TR], vob[TR], vxp[TR];
					const bool zt = zero_tile(j * TR);
#pragma unroll
					for(int u = 0; u < TR; ++u) {
						vlo[u] = tile[u * 32]; vdi[u] = tile[SLOT_DOUBLES + u * 32]; vup[u] = tile[2 * SLOT_DOUBLES + u * 32];
						vob[u] = american && !zt ? tile[3 * SLOT_DOUBLES + u * 32] : 0.;
						vx[u] = tile[4 * SLOT_DOUBLES + u * 32];
						vxp[u] = any_b2 ? tile[5 * SLOT_DOUBLES + u * 32] : 0.;
					}
					__syncwarp();
					if(j + stages < nb) { const int row_ = (j + stages) * TR; if(zero_tile(row_)) ring_issue(ring, maps, blk0, bz, row_, lane); else ring_issue(ring, maps, blk0, bx, row_, lane); }
#pragma unroll
					for(int u = 0; u < TR; ++u) {
						const int i = j * TR + u;
						if(i >= 1 && i <= N) finish_row(i - 1, vx[u]);
						xm = xc; xc = vx[u]; pxp = vxp[u]; plo = vlo[u]; pdi = vdi[u]; pup = vup[u]; pob = vob[u];
					}
				}
				if(nb * TR == N) finish_row(N - 1, 0.);
			} else {
			// ---- forward elimination of (lA + P) x = lb + P obstacle ------
			{
				// slots: lo di up | lb | obstacle x
				const Boxes bx = { { 3, 1, american ? 2 : 0 }, { ta.coef, ta.lb, ta.obx }, { 0, 3, 4 } };
				const Boxes bz = { { 3, 1, 1 }, { ta.coef, ta.lb, ta.x }, { 0, 3, 5 } };      // zero-obstacle tiles
				// rows before the first changed active-set bit keep their cp, dp (see the header)
				const int js = american ? fstart / TR : 0;
				const int lead2 = nb - js < stages ? nb - js : stages;
				pass_begin();
				for(int j = 0; j < lead2; ++j) { const int row_ = (js + j) * TR; if(zero_tile(row_)) ring_issue(ring, maps, blk0, bz, row_, lane); else ring_issue(ring, maps, blk0, bx, row_, lane); }
				cprev = 0.; dprev = 0.;
				if(js > 0) { cprev = cp[(size_t)(js * TR - 1) * ld]; dprev = dp[(size_t)(js * TR - 1) * ld]; }
				for(int j = js; j < nb; ++j) {
					const int s = ring_wait(ring);
					const double *tile = sm + s * STAGE_DOUBLES;
					double vlo[TR], vdi[TR], vup[TR], vlb[TR], vob[TR], vx[TR];
					const bool zt = zero_tile(j * TR);
#pragma unroll
					for(int u = 0; u < TR; ++u) {
						vlo[u] = tile[u * 32]; vdi[u] = tile[SLOT_DOUBLES + u * 32]; vup[u] = tile[2 * SLOT_DOUBLES + u * 32];
						vlb[u] = tile[3 * SLOT_DOUBLES + u * 32];
						vob[u] = american && !zt ? tile[4 * SLOT_DOUBLES + u * 32] : 0.;
						vx[u] = american ? tile[5 * SLOT_DOUBLES + u * 32] : 0.;
					}
					__syncwarp();
					if(j + stages < nb) { const int row_ = (j + stages) * TR; if(zero_tile(row_)) ring_issue(ring, maps, blk0, bz, row_, lane); else ring_issue(ring, maps, blk0, bx, row_, lane); }
#pragma unroll
					for(int u = 0; u < TR; ++u) {
						const int i = j * TR + u;
						if(i >= N) break;
						eliminate(i, vlo[u], vdi[u], vup[u], vlb[u], vob[u], vx[u], again);
					}
				}
			}
			}
			// ---- back substitution + relativeError --------------------------
			// With the penalty iteration the pass also reads the obstacle and compares the active
			// set of the new iterate with the one this solve was assembled with: if nothing
			// changed, the next iteration would assemble the same matrix and right-hand side and
			// return this iterate bit for bit (relativeError == 0) — it is counted, not computed.
			// A warp makes as many passes as its slowest contract needs, and nearly every step
			// has some contract whose free boundary has just crossed a node (three iterations
			// instead of two): without the shortcut every warp would take three solves per step.
			bool notdone = false, changed = false;
			int fc = N;
			{
				// slots: cp dp | obstacle x   (iterate == american in this kernel)
				const Boxes bx = { { 2, american ? 2 : 0, 0 }, { ta.cpdp, ta.obx, -1 }, { 0, 2, 0 } };
				const Boxes bz = { { 2, 1, 0 }, { ta.cpdp, ta.x, -1 }, { 0, 3, 0 } };      // zero-obstacle tiles
				pass_begin();
				for(int j = 0; j < lead; ++j) { const int row_ = N - (j + 1) * TR; if(zero_tile(row_)) ring_issue(ring, maps, blk0, bz, row_, lane); else ring_issue(ring, maps, blk0, bx, row_, lane); }
				double xn = 0.;
				for(int j = 0; j < nb; ++j) {
					const int r0 = N - (j + 1) * TR;      // first row of the tile (negative in the last one: those rows are skipped)
					const int s = ring_wait(ring);
					const double *tile = sm + s * STAGE_DOUBLES;
					double vcp[TR], vdp[TR], vxo[TR], vob[TR];
					const bool zt = zero_tile(r0);
#pragma unroll
					for(int u = 0; u < TR; ++u) {
						vcp[u] = tile[u * 32]; vdp[u] = tile[SLOT_DOUBLES + u * 32];
						vob[u] = american && !zt ? tile[2 * SLOT_DOUBLES + u * 32] : 0.;
						vxo[u] = iterate ? tile[3 * SLOT_DOUBLES + u * 32] : 0.;
					}
					__syncwarp();
					if(j + stages < nb) { const int row_ = N - (j + stages + 1) * TR; if(zero_tile(row_)) ring_issue(ring, maps, blk0, bz, row_, lane); else ring_issue(ring, maps, blk0, bx, row_, lane); }
#pragma unroll
					for(int u = TR - 1; u >= 0; --u) {
						const int i = r0 + u;
						if(i < 0) break;
						xn = vdp[u] - vcp[u] * xn;
						if(iterate) {
							const double xo = vxo[u];
							const double fa = fabs(xn), fb = fabs(xo);
							double d = fa < fb ? fb : fa;
							d = scale < d ? d : scale;
							// IterativeMethod.hpp:1131-1135: |a - b| / max(scale, |a|, |b|) > tol.  The quotient is
							// formed only in the rounding band around the threshold (never in practice); outside
							// it the product form gives the same decision without a division per row.
							const double diff = fabs(xn - xo), bound = tol * d;
							if(diff > bound * (1. + 8e-15)) notdone = true;
							else if(diff >= bound * (1. - 8e-15) && diff / d > tol) notdone = true;
							if(american && ((xn < vob[u]) != (xo < vob[u]))) { changed = true; fc = i; }   // descending: ends as the first changed row
						}
						if(again) x[(size_t)i * ld] = xn;
					}
				}
			}
			if(again) {
				++its; ++computed;
				again = iterate && notdone;
				if(again && !changed) {
					// same active set: the confirming iteration is counted (see above)
					if(its >= b.max_inner) status = 1; else ++its;
					again = false;
				} else if(again && its >= b.max_inner) { status = 1; again = false; }
			}
			fstart = __reduce_min_sync(FULL, again ? fc : N);      // warp-uniform; N when no contract iterates
		} while(__any_sync(FULL, again));

		if(act) {
			if(out.inner && n_steps < b.max_steps) {
				out.inner[(size_t)c * out.inner_stride + n_steps] = (uint8_t)(its > 255 ? 255 : its);
			}
			inner_total += its;
			++n_steps;
			time0 = time1; time1 = t;
			ns.end_step();

			if(variable && more) {
				// dt *= target / (relativeError(iterand(0), iterand(1), scale) + epsilon), Stepper.hpp:68-82
				// (once per step, ordinary loads: independent, so they stream)
				double err = 0.;
				for(int i = 0; i < N; ++i) {
					const double a = x[(size_t)i * ld], v = xp[(size_t)i * ld];
					const double fa = fabs(a), fb = fabs(v);
					double dn = fa < fb ? fb : fa;
					dn = b.variable_scale < dn ? dn : b.variable_scale;
					const double e = fabs(a - v) / dn;
					err = e > err ? e : err;
				}
				vdt *= vtarget / (err + DBL_EPSILON);
			}

			if(st.event_after) {
				// PenaltyMethod::constraintMask (PenaltyMethod.hpp:127-139) reads the tolerance iteration's
				// last iterand, which the events below do not touch: with user events the mask is taken
				// here, before them; the last step of a sweep always ends in an event
				if(out.active && ob && b.n_exercise > 0) {
					for(int i = 0; i < N; ++i) {
						const double pred = (0. + 1. * x[(size_t)i * ld]) - ob[(size_t)i * ld];
						out.active[(size_t)c * out.active_stride + i] = pred < 0. ? 1 : 0;
					}
				}
				// Events at one time go in descending add() order (IterativeMethod.hpp:1340-1352).
				// Rare (E times per sweep): ordinary loads and stores.
				for(int ue = 0; ue < st.user_events; ++ue) {
					for(int op = 0; op < 2; ++op) {
						const bool is_dividend = (op == 0) == (b.dividend_first != 0);
						if(is_dividend && b.dividend_kind != QPDE_DIVIDEND_NONE) {
							// Event<1>::_doEvent with V(max(S - D, 0)) or V((1 - D) S) through the
							// piecewise-linear interpolant (Core/Event.hpp:100-112, Interpolant.hpp:153-181,
							// 331-374).  The shifted price never exceeds S_i, so the nodes are updated in
							// place from the top down, and the bracketing interval only moves down.
							const double D = b.dividend[c];
							const double *S = w.S + base;
							const double Sfirst = S[0], Slast = S[(size_t)(N - 1) * ld];
							int idx = N - 2;
							for(int i = N - 1; i >= 0; --i) {
								const double Si = S[(size_t)i * ld];
								const double at = b.dividend_kind == QPDE_DIVIDEND_CASH ? (Si - D > 0. ? Si - D : 0.) : (1. - D) * Si;
								int k; double wgt;
								if(at <= Sfirst) { k = 0; wgt = 1.; }
								else if(at >= Slast) { k = N - 2; wgt = 0.; }
								else {
									while(idx > 0 && S[(size_t)idx * ld] > at) --idx;      // S[idx] <= at < S[idx + 1]
									k = idx;
									wgt = (S[(size_t)(k + 1) * ld] - at) / (S[(size_t)(k + 1) * ld] - S[(size_t)k * ld]);
								}
								double sum = 0.;
								if(wgt > DBL_EPSILON) sum += wgt * x[(size_t)k * ld];
								if(1. - wgt > DBL_EPSILON) sum += (1. - wgt) * x[(size_t)(k + 1) * ld];
								x[(size_t)i * ld] = sum;
							}
						}
						if(!is_dividend && evob) {
							for(int i = 0; i < N; ++i) {
								const size_t at = (size_t)i * ld;
								const double ex = evob[at], v = x[at];
								x[at] = v > ex ? v : ex;       // QuantPDE::max, EarlyInclude.hpp:48
							}
						}
					}
				}
				ns.after_event();
			}
		}
	}

	if(!valid) return;
	// ---- outputs (contract-major) ------------------------------------------
	if(out.V) for(int i = 0; i < N; ++i) out.V[(size_t)c * out.V_stride + i] = x[(size_t)i * ld];
	if(out.S) for(int i = 0; i < N; ++i) out.S[(size_t)c * out.S_stride + i] = w.S[(size_t)i * ld + base];
	if(out.active && !(ob && b.n_exercise > 0)) {
		for(int i = 0; i < N; ++i) {
			const double pred = ob ? (0. + 1. * x[(size_t)i * ld]) - ob[(size_t)i * ld] : 0.;
			out.active[(size_t)c * out.active_stride + i] = pred < 0. ? 1 : 0;
		}
	}
	if(out.value) {
		// PiecewiseLinear::interpolate (Core/Interpolant.hpp:153-181,331-374)
		const double *S = w.S + base;
		const double s0 = b.spot ? b.spot[c] : b.strike[c];
		int idx; double wgt;
		if(s0 <= S[0]) { idx = 0; wgt = 1.; }
		else if(s0 >= S[(size_t)(N - 1) * ld]) { idx = N - 2; wgt = 0.; }
		else {
			int l = 0, hgh = N - 2, mid = 0; wgt = 0.;
			while(l <= hgh) {
				mid = (l + hgh) / 2;
				if(s0 < S[(size_t)mid * ld]) hgh = mid - 1;
				else if(s0 >= S[(size_t)(mid + 1) * ld]) l = mid + 1;
				else {
					wgt = (S[(size_t)(mid + 1) * ld] - s0)
						/ (S[(size_t)(mid + 1) * ld] - S[(size_t)mid * ld]);
					break;
				}
			}
			idx = mid;
		}
		double sum = 0.;
		if(wgt > DBL_EPSILON) sum += wgt * x[(size_t)idx * ld];
		if(1. - wgt > DBL_EPSILON) sum += (1. - wgt) * x[(size_t)(idx + 1) * ld];
		out.value[c] = sum;
	}
	if(out.n_steps) out.n_steps[c] = n_steps;
	if(out.inner_total) out.inner_total[c] = inner_total;
	if(out.computed) out.computed[c] = computed;
	if(out.status) out.status[c] = status;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
	const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
	CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_entry() {
	static EncodeTiledFn fn = [] {
		void *p = nullptr;
		cudaDriverEntryPointQueryResult q;
		if(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess
				|| q != cudaDriverEntryPointSuccess) p = nullptr;
		return (EncodeTiledFn)p;
	}();
	return fn;
}

} // namespace

bool interleaved_tma_supported(const DeviceBatch &b) {
	// policy iteration walks several sets of operator rows and BDF3 .. BDF6 a ring of history
	// arrays: those batches stay on k_interleaved_sweep
	// With the confirming iteration counted and the right-hand side fused into the first solve the
	// ring also wins on small systems (measured at 65 / 129 / 257 nodes: 2.9 against 2.0 - 2.3 TB/s
	// algorithmic for the plain-load sweep); below a few tiles per pass it is all prologue.
	int min_nodes = 32;
	if(const char *e = std::getenv("QPDE_TMA_MIN_NODES")) { const int v = std::atoi(e); if(v >= 2) min_nodes = v; }   // tuning knob
	return b.n_controls == 0 && b.N >= min_nodes && b.scheme <= QPDE_SCHEME_BDF2 && !b.event_program && !b.source_table;   // event programs, operator sources: plain-load sweep
}

bool interleaved_tma_prepare(InterleavedTma *t, double *block, long long rows, int arrays, std::string *error) {
	static_assert(sizeof(CUtensorMap) <= sizeof(t->map[0]), "CUtensorMap does not fit InterleavedTma::map");
	t->enabled = false;
	EncodeTiledFn encode = encode_tiled_entry();
	if(!encode) { if(error) *error = "cuTensorMapEncodeTiled is not available from this driver"; return false; }
	const cuuint64_t dims[3] = { (cuuint64_t)QPDE_WARP_ROW, (cuuint64_t)rows, (cuuint64_t)arrays };
	const cuuint64_t strides[2] = { (cuuint64_t)QPDE_WARP_ROW * sizeof(double), (cuuint64_t)rows * QPDE_WARP_ROW * sizeof(double) };
	const cuuint32_t estr[3] = { 1u, 1u, 1u };
	for(int k = 1; k <= 3; ++k) {
		const cuuint32_t box[3] = { 32u, (cuuint32_t)TR, (cuuint32_t)k };
		const CUresult rc = encode(reinterpret_cast<CUtensorMap *>(t->map[k - 1]), CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, block,
			dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
			CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
		if(rc != CUDA_SUCCESS) {
			if(error) *error = "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)rc);
			return false;
		}
	}
	t->stages = 0;   // chosen per launch
	t->enabled = true;
	return true;
}

// Ring depth.  A stage is 6 KB per warp; what hides the HBM latency is the number of bytes in
// flight per SM, (warps per SM) x stages x 6 KB.  Up to seven warps per SM the whole batch is
// resident at once and the shared memory is split between the warps there are; beyond that,
// seven warps with five stages each (fewer, deeper rings would add a wave of CTAs).
static int tma_stages(int C, int sm_count) {
	if(const char *e = std::getenv("QPDE_TMA_STAGES")) { const int v = std::atoi(e); if(v >= 1 && v <= 32) return v; }
	const int warps = (C + 31) / 32;
	int per_sm = (warps + sm_count - 1) / sm_count;
	if(per_sm > 7) per_sm = 7;
	int stages = (225 * 1024 / per_sm - 1024) / (STAGE_BYTES + 8);
	return stages > 16 ? 16 : (stages < 2 ? 2 : stages);
}

void launch_interleaved_tma(const DeviceBatch &b, const InterleavedWork &w, const InterleavedTma &t, double *block,
		const DeviceResult &out, cudaStream_t stream, long long *launches) {
	launch_interleaved_setup(b, w, stream, launches);
	// positions in the stack of arrays (qpde_api.cu lays it out S, lo, di, up, lb, [obstacle], x, cp, dp, ...)
	const long long per_array = (long long)b.N * w.ld;
	auto index_of = [&](const double *p) { return p ? (int)((p - block) / per_array) : -1; };
	TmaArrays ta;
	ta.coef = index_of(w.lo); ta.lb = index_of(w.lb); ta.obx = index_of(w.obst); ta.x = index_of(w.x);
	ta.cpdp = index_of(w.cp); ta.xp = index_of(w.xprev);
	int device = 0, sm_count = 148;
	cudaGetDevice(&device);
	cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, device);
	const int stages = t.stages > 0 ? t.stages : tma_stages(b.C, sm_count);
	const int smem = stages * STAGE_BYTES + stages * 8;
	const int blocks = (b.C + 31) / 32;
	TmaMaps maps;
	for(int k = 0; k < 3; ++k) maps.k[k] = *reinterpret_cast<const CUtensorMap *>(t.map[k]);
	if(b.american) {
		cudaFuncSetAttribute(k_interleaved_sweep_tma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
		k_interleaved_sweep_tma<true><<<blocks, 32, smem, stream>>>(maps, b, w, out, ta, stages);
	} else {
		cudaFuncSetAttribute(k_interleaved_sweep_tma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
		k_interleaved_sweep_tma<false><<<blocks, 32, smem, stream>>>(maps, b, w, out, ta, stages);
	}
	*launches += 1;
}

} // namespace qpde
