// kernels_impl.cuh -- sm_100a device code of the level-0 LBM time step, templated on the lattice (compiled
// with -fmad=false).  Included by one translation unit per lattice (kernels_d2q9.cu, kernels_d3q19.cu,
// kernels_d3q27.cu: they build in parallel) and by kernels_common.cu for the layout-conversion templates.
//
// One step = k_step over all planes (fluid sites) + k_bc over the list of velocity/pressure sites.
// Both read lattice `fin` and write lattice `fout` (two-lattice pull scheme, SoA populations), so
// they are independent of each other and of the order sites are visited in; the reference's
// loop-order dependent "update the neighbour on the fly" (optimised.cpp:1375-1404) is reproduced by
// recomputing the neighbour's stream+macro inside the boundary thread from `fin`.
//
// Reference for every function: /root/reference/LUMA/src/GridObj_ops_lbm_optimised.cpp (cited as
// optimised.cpp below).
#pragma once
#include "kernels.cuh"

namespace luma {

// tuning knobs (defaults are the measured best, profiles/r01_variants.txt)
#ifndef LUMA_STEP_THREADS
#define LUMA_STEP_THREADS 128
#endif
#ifndef LUMA_MIN_BLOCKS
#define LUMA_MIN_BLOCKS 6
#endif
#ifndef LUMA_MIN_BLOCKS_SMAG
#define LUMA_MIN_BLOCKS_SMAG 6  /* with LUMA_SMAG_RECOMPUTE the Smagorinsky kernel fits 80 registers: 6 x 128 threads like BGK
                                   (measured: +7 % over 5 CTAs x 96 registers with the equilibrium kept, profiles/r02_probe_smag.txt) */
#endif
#ifndef LUMA_MIN_BLOCKS_KBC
#define LUMA_MIN_BLOCKS_KBC 3   /* KBC keeps the own-site populations, ds and dh alive beside feq */
#endif
#ifndef LUMA_VARIANTS
#define LUMA_VARIANTS 0         /* 1: also build the two measured-and-retired forms of k_step (k_step_v2: two sites per thread with 128-bit
                                   accesses, LUMA_B200_V2=1; k_step_tma: loads staged through shared memory by the TMA engine,
                                   LUMA_B200_TMA=1) -- luma_b200.build.build_variant("variants", ["-DLUMA_VARIANTS=1"]); profiles/r02_variants.txt */
#endif
#ifndef LUMA_SMAG_RECOMPUTE
#define LUMA_SMAG_RECOMPUTE 1   /* k_step's Smagorinsky variant evaluates the equilibrium twice instead of keeping it in registers
                                   (0 + LUMA_MIN_BLOCKS_SMAG=5: the round-1 form) */
#endif
#ifndef LUMA_LOAD_MODE
#define LUMA_LOAD_MODE 0      /* 0 ld.global.nc (__ldg), 1 ld.global.cs, 2 ld.global.nc.L1::no_allocate, 3 plain */
#endif
#ifndef LUMA_STORE_MODE
#define LUMA_STORE_MODE 0     /* 0 plain, 1 st.global.cs, 2 st.global.L1::no_allocate */
#endif
constexpr int STEP_THREADS = LUMA_STEP_THREADS;
// collision operator of a kernel instantiation: L_USE_BGKSMAG / L_USE_KBC_COLLISION (optimised.cpp:147-151, :769-773)
enum : int { COLL_BGK = 0, COLL_SMAG = 1, COLL_KBC = 2 };
template <class L, int COLL> constexpr int step_min_blocks()
{
	return COLL == COLL_SMAG ? LUMA_MIN_BLOCKS_SMAG : (COLL == COLL_KBC ? (L::Q == 27 ? 2 : LUMA_MIN_BLOCKS_KBC) : LUMA_MIN_BLOCKS);
}

__device__ __forceinline__ double load_pop(const double *p)
{
#if LUMA_LOAD_MODE == 0
	return __ldg(p);
#elif LUMA_LOAD_MODE == 1
	return __ldcs(p);
#elif LUMA_LOAD_MODE == 2
	double v;
	asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
	return v;
#else
	return *p;
#endif
}
__device__ __forceinline__ void store_pop(double *p, double v)
{
#if LUMA_STORE_MODE == 0
	*p = v;
#elif LUMA_STORE_MODE == 1
	__stcs(p, v);
#else
	asm volatile("st.global.L1::no_allocate.f64 [%0], %1;" :: "l"(p), "d"(v) : "memory");
#endif
}

// ------------------------------------------------------------------------------------------------
// pull-stream of one site, GridObj::_LBM_stream_opt (optimised.cpp:206-297): population v comes
// from site - c_v (periodic by wrap, :216-218) unless that site is eSolid, in which case the
// site's own opposite population bounces back (:238-243).  All 19 loads are independent.
// ------------------------------------------------------------------------------------------------
template <class L>
__device__ __forceinline__ void pull_populations(const StepArgs &a, const int p, const unsigned r, const long long id,
	const uint32_t w, double (&f)[L::Q])
{
	const double *base = a.fin + id;
	const bool x_wraps = a.wrap_x && (p == 0 || p == a.P - 1);
	if ((w & (CW<L>::LINKS | CW<L>::EDGE)) == 0 && !x_wraps)
	{
		// interior site with fluid neighbours only (the overwhelmingly common case): kernel-uniform offsets
		const char *pb = reinterpret_cast<const char *>(base);
#pragma unroll
		for (int v = 0; v < L::Q; ++v) f[v] = load_pop(reinterpret_cast<const double *>(pb + a.off_pull[v]));
		return;
	}
	long long xm = -(long long)a.MK, xp = (long long)a.MK;      // offsets to x-1 / x+1
	if (a.wrap_x)
	{
		if (p == 0) xm = (long long)(a.P - 1) * a.MK;
		if (p == a.P - 1) xp = -(long long)(a.P - 1) * a.MK;
	}
	long long ym = -(long long)a.K, yp = (long long)a.K, zm = -1, zp = 1;
	if (w & CW<L>::EDGE)
	{
		const unsigned j = r / (unsigned)a.K, k = r - j * (unsigned)a.K;
		if (j == 0) ym = (long long)(a.M - 1) * a.K;
		if (j == (unsigned)a.M - 1) yp = -(long long)(a.M - 1) * a.K;
		if (L::D == 3)
		{
			if (k == 0) zm = a.K - 1;
			if (k == (unsigned)a.K - 1) zp = -(long long)(a.K - 1);
		}
	}
#pragma unroll
	for (int v = 0; v < L::Q; ++v)
	{
		const int cx = L::c(v, 0), cy = L::c(v, 1), cz = L::c(v, 2);
		long long off = (long long)v * a.stride;
		if (cx == 1) off += xm; else if (cx == -1) off += xp;
		if (cy == 1) off += ym; else if (cy == -1) off += yp;
		if (cz == 1) off += zm; else if (cz == -1) off += zp;
		if (v < L::Q - 1)
		{
			const long long off_bb = (long long)opposite<L>(v) * a.stride;
			if ((w >> v) & 1u) off = off_bb;
		}
		f[v] = load_pop(base + off);
	}
}

// the populations of the previous time level at the site itself (what the reference's KBC operator collides)
template <class L>
__device__ __forceinline__ void load_own(const StepArgs &a, const long long id, double (&fo)[L::Q])
{
	const char *pb = reinterpret_cast<const char *>(a.fin + id);
	const long long sb = a.stride * (long long)sizeof(double);
#pragma unroll
	for (int v = 0; v < L::Q; ++v) fo[v] = load_pop(reinterpret_cast<const double *>(pb + (long long)v * sb));
}

// The pull of a whole warp of step_site, arranged so that a wall next to the warp does not split it: when no lane needs a
// periodic wrap, every lane -- plain fluid sites, sites with bounce-back links, and the never-updated sites that only complete
// a sector (`through`) -- takes ONE load sequence whose per-population offset is selected without a branch:
//     pulled population (kernel-uniform offset)  |  own opposite population (bounce-back)  |  own population (pass-through).
// Warps without links or pass-through lanes keep the pure kernel-uniform sequence; lanes on the first/last row or column
// (periodic wrap) or on a wrapping x-plane use the general pull_populations.  (`mode` is warp-uniform: 0 pure, 1 select.)
template <class L>
__device__ __forceinline__ void pull_warp(const StepArgs &a, const int p, const unsigned r, const long long id, const uint32_t w,
	const bool fluid, const bool through, const int mode, double (&f)[L::Q])
{
	const bool x_wraps = a.wrap_x && (p == 0 || p == a.P - 1);
	if (mode == 0 || (w & CW<L>::EDGE) != 0 || x_wraps)
	{
		if (fluid) pull_populations<L>(a, p, r, id, w, f);
		else load_own<L>(a, id, f);
		return;
	}
	const char *pb = reinterpret_cast<const char *>(a.fin + id);
	const long long sb = a.stride * (long long)sizeof(double);
#pragma unroll
	for (int v = 0; v < L::Q; ++v)
	{
		long long off = a.off_pull[v];
		if (v < L::Q - 1 && ((w >> v) & 1u)) off = (long long)opposite<L>(v) * sb;
		if (through) off = (long long)v * sb;
		f[v] = load_pop(reinterpret_cast<const double *>(pb + off));
	}
}

// BGK(/Smagorinsky) collision with optional Guo forcing, GridObj::_LBM_collide_opt (optimised.cpp:765-790),
// or the KBC operator _LBM_kbcCollide_opt (:1122-1305), which replaces f by the collided own-site populations
// f += omega_s * (feq - f) (+ Guo force), the relaxation of GridObj::_LBM_collide_opt (optimised.cpp:775-787)
template <class L, int FORCE>
__device__ __forceinline__ void relax(const StepArgs &a, const double (&u)[3], const double (&feq)[L::Q], const double omega_s, double (&f)[L::Q])
{
#pragma unroll
	for (int v = 0; v < L::Q; ++v)
	{
		if constexpr (FORCE != 0)
			f[v] = f[v] + (omega_s * (feq[v] - f[v]) + guo_force<L, (FORCE > 0 ? FORCE - 1 : 0)>(v, u, a.Fg, a.C, a.lam));
		else
			f[v] = f[v] + omega_s * (feq[v] - f[v]);
	}
}

template <class L, int COLL, int FORCE>
__device__ __forceinline__ void collide(const StepArgs &a, const long long id, const double (&u)[3], const double (&feq)[L::Q], double (&f)[L::Q])
{
	if constexpr (COLL == COLL_KBC)
	{
		double fo[L::Q];
		load_own<L>(a, id, fo);
		kbc_collide<L, FORCE>(u, feq, fo, a.kbc_beta_m1, a.kbc_inv_beta, a.Fg, a.C, a.lam, f);
		return;
	}
	double omega_s = a.omega;
	if (COLL == COLL_SMAG) omega_s = smagorinsky_omega<L>(f, feq, a.tau, a.smag_coef);
	relax<L, FORCE>(a, u, feq, omega_s, f);
}

template <class L>
__device__ __forceinline__ void store_populations(const StepArgs &a, const long long id, const double (&f)[L::Q])
{
	char *pb = reinterpret_cast<char *>(a.fout + id);
	const long long sb = a.stride * (long long)sizeof(double);
#pragma unroll
	for (int v = 0; v < L::Q; ++v) store_pop(reinterpret_cast<double *>(pb + (long long)v * sb), f[v]);
}

// Fused halo exchange: a site on one of the slab's two face planes also stores the populations that leave the slab
// straight into the neighbour GPU's ghost plane (peer memory over NVLink) -- c_x = -1 from the first owned plane into
// the left neighbour's high ghost plane, c_x = +1 from the last owned plane into the right neighbour's low ghost
// plane: what MpiManager::mpi_communicate packs, sends and unpacks (src/MpiManager.cpp:631-815) becomes part of the
// producing kernel's epilogue.  Arrival is signalled afterwards by k_halo_publish.
template <class L>
__device__ __forceinline__ void store_outgoing(const StepArgs &a, const int p, const unsigned r, const double (&f)[L::Q])
{
	if (p == 1 && a.peer_f[0])
	{
		double *dst = a.peer_f[0] + (long long)(a.peer_P[0] - 1) * a.MK + r;
#pragma unroll
		for (int v = 0; v < L::Q; ++v)
			if (L::c(v, 0) == -1) dst[(long long)v * a.peer_stride[0]] = f[v];
	}
	if (p == a.P - 2 && a.peer_f[1])
	{
		double *dst = a.peer_f[1] + r;
#pragma unroll
		for (int v = 0; v < L::Q; ++v)
			if (L::c(v, 0) == 1) dst[(long long)v * a.peer_stride[1]] = f[v];
	}
}

// ------------------------------------------------------------------------------------------------
// per-link stream of a site that is, or pulls from, one of the special types -- the full branch
// ladder of GridObj::_LBM_stream_opt (optimised.cpp:206-297) evaluated from the eType array:
// specular reflection on eSlip sites (_LBM_applySpecReflect_opt :527-581: the first direction, in the
// order x, y, z, in which the link's component equals the inward normal's reflects the link),
// halfway bounce-back (:238-243), eExtrapolateRight (:246-251: the same population two planes to the
// left of the source), forced-equilibrium eVelocity (:254-270, non-regularised builds only), copy.
// ------------------------------------------------------------------------------------------------
template <class L>
__device__ __noinline__ void pull_general(const StepArgs &a, const int p, const int j, const int k, const long long id,
	const uint32_t desc, const uint8_t type, double (&f)[L::Q])
{
	int n[3];
#pragma unroll
	for (int d = 0; d < 3; ++d) n[d] = (int)((desc >> (CW_N_SHIFT + 2 * d)) & 3u) - 1;
	const double *fin = a.fin;
#pragma unroll
	for (int v = 0; v < L::Q; ++v)
	{
		const int cx = L::c(v, 0), cy = L::c(v, 1), cz = L::c(v, 2);
		if (type == T_SLIP)
		{
			if (cx != 0 && n[0] == cx) { f[v] = fin[(long long)reflect<L>(v, 0) * a.stride + id]; continue; }
			if (cy != 0 && n[1] == cy) { f[v] = fin[(long long)reflect<L>(v, 1) * a.stride + id]; continue; }
			if (L::D == 3 && cz != 0 && n[2] == cz) { f[v] = fin[(long long)reflect<L>(v, 2) * a.stride + id]; continue; }
		}
		int sp = p - cx, sj = j - cy, sk = k - cz;
		if (a.wrap_x) { if (sp < 0) sp += a.P; else if (sp >= a.P) sp -= a.P; }
		if (sj < 0) sj += a.M; else if (sj >= a.M) sj -= a.M;
		if (sk < 0) sk += a.K; else if (sk >= a.K) sk -= a.K;
		const long long src = ((long long)sp * a.M + sj) * a.K + sk;
		const uint8_t st = a.types[src];
		if (st == T_SOLID)
			f[v] = fin[(long long)opposite<L>(v) * a.stride + id];
		else if (st == T_EXTRAPOLATE_RIGHT)
			f[v] = fin[(long long)v * a.stride + src - 2 * (long long)a.MK];
		else if (!a.regularised && st == T_VELOCITY)
		{
			// u of the source site is u_in[j of THIS site] * ramp(t*dt) when a ramp is defined (:258-263)
			double us[3] = { 0.0, 0.0, 0.0 }, feq[L::Q];
#pragma unroll
			for (int d = 0; d < L::D; ++d)
				us[d] = a.velramp_on ? a.uin[d * a.M + j] * a.ramp_t : a.u[(long long)d * a.stride + src];
			equilibrium_all<L>(a.rho[src], us, a.C, feq);
			f[v] = feq[v];
		}
		else
			f[v] = fin[(long long)v * a.stride + src];
	}
}

// wall descriptor of a site (normalDirection, normal vector, edgeCount at bits 22..31): part of the cell word
// where it fits, else the per-site descriptor array
template <class L>
__device__ __forceinline__ uint32_t site_desc(const StepArgs &a, const long long id, const uint32_t w)
{
	if constexpr (L::DESC_IN_WORD) return w;
	else return a.bcdesc[id];
}

// stream of a site handled by k_bc: the fast path unless the grid holds special types
template <class L>
__device__ __forceinline__ void pull_any(const StepArgs &a, const int p, const unsigned r, const int j, const int k,
	const long long id, const uint32_t w, double (&f)[L::Q])
{
	if (a.general && cw_class<L>(w) != CLS_FLUID) pull_general<L>(a, p, j, k, id, site_desc<L>(a, id, w), a.types[id], f);
	else pull_populations<L>(a, p, r, id, w, f);
}

// ------------------------------------------------------------------------------------------------
// time-averaged statistics, the tail of GridObj::_LBM_macro_opt (optimised.cpp:895-917):
//   avg = (avg * (double)t + x) / (double)(t + 1)   for rho, u_p and u_p*u_q (p <= q).
// `reps` > 1 reproduces the reference advancing the averages of a site once more every time a
// boundary site "updates it on the fly" (_LBM_updateInteriorLatticeSite :1422-1434 calls
// _LBM_macro_opt, which ends with this block).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double tavg_advance(double avg, const double x, const double t_now, const double t_next, const int reps)
{
	for (int r = 0; r < reps; ++r)
	{
		double ta = avg * t_now;
		ta = ta + x;
		avg = ta / t_next;
	}
	return avg;
}

template <class L>
__device__ __forceinline__ void tavg_update(const StepArgs &a, const long long id, const double rho, const double (&u)[3], const int reps)
{
	double *t = a.tav + id;
	t[0] = tavg_advance(t[0], rho, a.t_now, a.t_next, reps);
	int pq = 0;
#pragma unroll
	for (int p = 0; p < L::D; ++p)
	{
		double *tp = t + (long long)(1 + p) * a.stride;
		*tp = tavg_advance(*tp, u[p], a.t_now, a.t_next, reps);
#pragma unroll
		for (int q = p; q < L::D; ++q)
		{
			double *tq = t + (long long)(1 + L::D + pq) * a.stride;
			*tq = tavg_advance(*tq, u[p] * u[q], a.t_now, a.t_next, reps);
			++pq;
		}
	}
}

// ------------------------------------------------------------------------------------------------
// the hot kernel: one thread per site of one x-plane; fluid sites only (optimised.cpp:91-156)
// ------------------------------------------------------------------------------------------------
// macro -> (time averages) -> equilibrium -> collide of one fluid site whose populations have been pulled into f;
// leaves the post-collision populations in f and returns rho, u (optimised.cpp:122-156)
template <class L, int COLL, int FORCE, bool TAVG>
__device__ __forceinline__ void update_site(const StepArgs &a, const long long id, double (&f)[L::Q], double &rho, double (&u)[3])
{
	macroscopic<L, FORCE>(f, a.hFg, rho, u);
	if (TAVG) tavg_update<L>(a, id, rho, u, 1);
#if LUMA_SMAG_RECOMPUTE
	if constexpr (COLL == COLL_SMAG)
	{
		// Smagorinsky needs the equilibrium twice -- for the non-equilibrium stress that gives omega_s, then for the relaxation --
		// and keeping it would hold f AND feq (76 registers) live across the sqrt / division chain.  It is evaluated twice instead
		// (the same expression on the same operands: identical bits); the empty asm keeps the compiler from merging the two.
		double omega_s;
		{
			double feq1[L::Q];
			equilibrium_all<L>(rho, u, a.C, feq1);
			omega_s = smagorinsky_omega<L>(f, feq1, a.tau, a.smag_coef);
		}
		double r2 = rho, u2[3] = { u[0], u[1], u[2] };
#ifdef __CUDACC__
		asm volatile("" : "+d"(r2), "+d"(u2[0]), "+d"(u2[1]), "+d"(u2[2]), "+d"(omega_s));
#endif
		double feq2[L::Q];
		equilibrium_all<L>(r2, u2, a.C, feq2);
		relax<L, FORCE>(a, u2, feq2, omega_s, f);
		return;
	}
#endif
	double feq[L::Q];
	equilibrium_all<L>(rho, u, a.C, feq);
	collide<L, COLL, FORCE>(a, id, u, feq, f);
}

// Pass-through of never-updated sites (eSolid, eRefined, non-regularised eVelocity) that share a 32-byte sector with a site
// this kernel updates -- the solid site at each end of a wall-bounded z-row: k = 0 solid, k = 1, 2, 3 fluid.  Without it the
// warp's store covers 24 of the sector's 32 bytes, and the L2 has to fetch the sector from DRAM to rebuild its ECC (ncu:
// lts__t_sectors_data_ecc = 2 sectors x rows x Q per step, 7-9 % of the step on walls in z, profiles/r02_probe_walls.txt).
// With it the solid site's lane joins the SAME store instructions with its own populations (fout = fin at the site itself), and
// every sector is written whole.  Semantically a no-op: both lattices hold the same populations at such sites for the whole
// run (optimised.cpp:91-95 skips them and f.swap(fNew) :159 exchanges two arrays that agree there).
template <class L, int COLL, int FORCE, bool TAVG, bool PEER>
__device__ __forceinline__ void step_site(const StepArgs &a)
{
	const unsigned r = blockIdx.x * STEP_THREADS + threadIdx.x;
	const bool inside = r < a.MK;
	const int p = a.p0 + (int)blockIdx.y * a.pstep;
	const long long id = (long long)p * a.MK + r;
	const uint32_t w = inside ? __ldg(a.cw + id) : 0u;
	const bool fluid = inside && cw_class<L>(w) == CLS_FLUID;
	bool through = false;
	int mode = 0;
#ifdef __CUDACC__
	if (a.fill_holes)
	{
		const unsigned updated = __ballot_sync(0xffffffffu, fluid);
		through = inside && !fluid && w == 0u && ((updated >> (threadIdx.x & 28u)) & 0xfu) != 0u;
		// warp-uniform: does any lane have a bounce-back link or complete a sector?  then all lanes take the select sequence
		mode = __any_sync(0xffffffffu, through || (fluid && (w & CW<L>::LINKS) != 0)) ? 1 : 0;
	}
#endif
	if (!fluid && !through) return;
	if (a.rest_only)
	{
		// second launch of the two-sites-per-thread variant: only the sites whose pair (r & ~1, r | 1) k_step_v2 left alone
		const uint32_t wp = __ldg(a.cw + (id ^ 1));
		const bool x_wraps = a.wrap_x && (p == 0 || p == a.P - 1);
		if (!fluid) return;
		if (!x_wraps && (w & (CW<L>::LINKS | CW<L>::EDGE)) == 0 && cw_class<L>(wp) == CLS_FLUID && (wp & (CW<L>::LINKS | CW<L>::EDGE)) == 0) return;
	}
	double f[L::Q], u[3], rho;
	pull_warp<L>(a, p, r, id, w, fluid, through, mode, f);
	if (fluid) update_site<L, COLL, FORCE, TAVG>(a, id, f, rho, u);
	store_populations<L>(a, id, f);
	if (!fluid) return;
	if (PEER) store_outgoing<L>(a, p, r, f);
	if (a.write_macro)
	{
		a.rho[id] = rho;
#pragma unroll
		for (int d = 0; d < L::D; ++d) a.u[(long long)d * a.stride + id] = u[d];
	}
}

#if LUMA_VARIANTS
// ------------------------------------------------------------------------------------------------
// Variant of k_step with TWO z-adjacent sites per thread and 128-bit accesses (LUMA_B200_V2=1; measured against the
// one-site kernel in profiles/r02_variants.txt).  Sites (r, r+1), r even: all Q stores and the loads of the populations
// with c_z = 0 are aligned 16-byte accesses (STG.128 / LDG.128); the populations with c_z = +-1 are pulled from an odd
// element offset, where a 16-byte access would be misaligned, and stay two 8-byte loads.  A pair takes this path only if
// both sites are plain fluid sites away from walls and array edges; everything else is left to a second launch of the
// one-site kernel (StepArgs::rest_only).
// ------------------------------------------------------------------------------------------------
#ifndef LUMA_MIN_BLOCKS_V2
#define LUMA_MIN_BLOCKS_V2 3
#endif
template <class L, int COLL, int FORCE, bool TAVG>
__global__ void __launch_bounds__(STEP_THREADS, LUMA_MIN_BLOCKS_V2) k_step_v2(const StepArgs a)
{
#ifdef __CUDACC__
	const unsigned r = 2u * (blockIdx.x * STEP_THREADS + threadIdx.x);
	if (r >= a.MK) return;
	const int p = a.p0 + (int)blockIdx.y * a.pstep;
	const long long id = (long long)p * a.MK + r;
	const uint2 ww = __ldg(reinterpret_cast<const uint2 *>(a.cw + id));
	const bool x_wraps = a.wrap_x && (p == 0 || p == a.P - 1);
	const bool plain0 = cw_class<L>(ww.x) == CLS_FLUID && (ww.x & (CW<L>::LINKS | CW<L>::EDGE)) == 0;
	const bool plain1 = cw_class<L>(ww.y) == CLS_FLUID && (ww.y & (CW<L>::LINKS | CW<L>::EDGE)) == 0;
	if (plain0 && plain1 && !x_wraps)
	{
		double fa[L::Q], fb[L::Q], ua[3], ub[3], rhoa, rhob;
		const char *pb = reinterpret_cast<const char *>(a.fin + id);
#pragma unroll
		for (int v = 0; v < L::Q; ++v)
		{
			if (L::c(v, 2) == 0)
			{
				const double2 t = __ldg(reinterpret_cast<const double2 *>(pb + a.off_pull[v]));
				fa[v] = t.x; fb[v] = t.y;
			}
			else
			{
				fa[v] = load_pop(reinterpret_cast<const double *>(pb + a.off_pull[v]));
				fb[v] = load_pop(reinterpret_cast<const double *>(pb + a.off_pull[v] + 8));
			}
		}
		update_site<L, COLL, FORCE, TAVG>(a, id, fa, rhoa, ua);
		update_site<L, COLL, FORCE, TAVG>(a, id + 1, fb, rhob, ub);
		char *po = reinterpret_cast<char *>(a.fout + id);
		const long long sb = a.stride * (long long)sizeof(double);
#pragma unroll
		for (int v = 0; v < L::Q; ++v) *reinterpret_cast<double2 *>(po + (long long)v * sb) = make_double2(fa[v], fb[v]);
		if (a.write_macro)
		{
			*reinterpret_cast<double2 *>(a.rho + id) = make_double2(rhoa, rhob);
#pragma unroll
			for (int d = 0; d < L::D; ++d) *reinterpret_cast<double2 *>(a.u + (long long)d * a.stride + id) = make_double2(ua[d], ub[d]);
		}
	}
#endif
}

#endif  // LUMA_VARIANTS

template <class L, int COLL, int FORCE, bool TAVG>
__global__ void __launch_bounds__(STEP_THREADS, step_min_blocks<L, COLL>()) k_step(const StepArgs a)
{
	step_site<L, COLL, FORCE, TAVG, false>(a);
}

// the same on the two face planes of a slab, with the outgoing populations also stored into the neighbours' ghost
// planes (fused halo exchange; two planes per step, so its occupancy is irrelevant)
template <class L, int COLL, int FORCE, bool TAVG>
__global__ void __launch_bounds__(STEP_THREADS) k_step_faces(const StepArgs a)
{
	step_site<L, COLL, FORCE, TAVG, true>(a);
}

#if LUMA_VARIANTS
// ------------------------------------------------------------------------------------------------
// Variant of k_step whose loads are STAGED THROUGH SHARED MEMORY BY THE TMA ENGINE (LUMA_B200_TMA=1; measured against
// the per-thread-load kernel in profiles/r02_variants.txt).  One CTA = one run of STEP_THREADS consecutive sites of an
// x-plane.  In the SoA layout the values those sites pull for population v are one CONTIGUOUS run of the lattice, shifted
// by c_v -- so thread 0 issues Q one-dimensional bulk copies (cp.async.bulk.shared.global, SASS UBLKCP) of ~1 KB each,
// all completing on one mbarrier, and every thread then reads its Q values from shared memory (conflict-free).  Bulk
// copies need 16-byte aligned addresses and sizes: a run shifted by an odd number of elements (c_z = +-1) is copied from
// one element earlier, two elements longer.  What TMA cannot express stays per thread: sites with a bounce-back link, on
// the first/last row or column (periodic wrap), or in a tile whose shifted runs leave the array (first/last planes) take
// the ordinary pull_populations path; stores are per thread as well, because a tile also holds sites this kernel must
// not write (solid sites keep their state, boundary sites belong to k_bc).
// ------------------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
#endif

template <class L, int COLL, int FORCE, bool TAVG>
__global__ void __launch_bounds__(STEP_THREADS, step_min_blocks<L, COLL>()) k_step_tma(const StepArgs a)
{
#ifdef __CUDACC__
	constexpr int T = STEP_THREADS;
	constexpr int ROW = T + 2;
	__shared__ __align__(16) double tile[L::Q][ROW];
	__shared__ __align__(8) unsigned long long bar;
	__shared__ int staged;
	const unsigned r0 = blockIdx.x * T;
	const int p = a.p0 + (int)blockIdx.y * a.pstep;
	const long long id0 = (long long)p * a.MK + r0;
	const int n = (int)((a.MK - r0) < (unsigned)T ? (a.MK - r0) : (unsigned)T);
	const long long cells = (long long)a.P * a.MK;
	if (threadIdx.x == 0)
	{
		// every shifted run must lie inside the lattice array, else the whole tile takes the per-thread path
		bool ok = !(a.wrap_x && (p == 0 || p == a.P - 1));
#pragma unroll
		for (int v = 0; v < L::Q; ++v)
		{
			const long long start = id0 - ((long long)L::c(v, 0) * a.MK + (long long)L::c(v, 1) * a.K + L::c(v, 2));
			const long long odd = start & 1;
			if (start - odd < 0 || start - odd + ((n + odd + 1) & ~1LL) > cells) ok = false;
		}
		staged = ok ? 1 : 0;
		const uint32_t b = smem_u32(&bar);
		asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(b));
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		if (ok)
		{
			uint32_t total = 0;
#pragma unroll
			for (int v = 0; v < L::Q; ++v)
			{
				const long long start = id0 - ((long long)L::c(v, 0) * a.MK + (long long)L::c(v, 1) * a.K + L::c(v, 2));
				total += (uint32_t)(((n + (start & 1) + 1) & ~1LL) * 8);
			}
			asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(b), "r"(total) : "memory");
#pragma unroll
			for (int v = 0; v < L::Q; ++v)
			{
				const long long start = id0 - ((long long)L::c(v, 0) * a.MK + (long long)L::c(v, 1) * a.K + L::c(v, 2));
				const long long odd = start & 1;
				const uint32_t bytes = (uint32_t)(((n + odd + 1) & ~1LL) * 8);
				const double *src = a.fin + (long long)v * a.stride + (start - odd);
				asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
					:: "r"(smem_u32(&tile[v][0])), "l"(src), "r"(bytes), "r"(b) : "memory");
			}
		}
	}
	__syncthreads();
	const unsigned r = r0 + threadIdx.x;
	const bool live = (int)threadIdx.x < n;
	const long long id = id0 + threadIdx.x;
	const uint32_t w = live ? __ldg(a.cw + id) : 0u;
	if (staged)
	{
		const uint32_t b = smem_u32(&bar);
		asm volatile("{\n .reg .pred q;\n LUMA_TMA_WAIT:\n mbarrier.try_wait.parity.shared::cta.b64 q, [%0], 0;\n @q bra LUMA_TMA_DONE;\n bra LUMA_TMA_WAIT;\n LUMA_TMA_DONE:\n}"
			:: "r"(b) : "memory");
	}
	if (!live || cw_class<L>(w) != CLS_FLUID) return;

	double f[L::Q], u[3], rho;
	if (staged && (w & (CW<L>::LINKS | CW<L>::EDGE)) == 0)
	{
		// parity of the shifted start == parity of c_z when M*K and K are even (checked by the launcher)
#pragma unroll
		for (int v = 0; v < L::Q; ++v) f[v] = tile[v][threadIdx.x + ((L::c(v, 2) != 0) ? 1 : 0)];
	}
	else pull_populations<L>(a, p, r, id, w, f);
	update_site<L, COLL, FORCE, TAVG>(a, id, f, rho, u);
	store_populations<L>(a, id, f);
	if (a.write_macro)
	{
		a.rho[id] = rho;
#pragma unroll
		for (int d = 0; d < L::D; ++d) a.u[(long long)d * a.stride + id] = u[d];
	}
#endif
}

#endif  // LUMA_VARIANTS

// new-time rho,u of an extrapolation neighbour: GridObj::_LBM_updateAndExtrapolate +
// _LBM_updateInteriorLatticeSite (optimised.cpp:1353-1434).  A fluid neighbour is streamed and
// "macro'd" from the old lattice here; any other type keeps its stored values (its macro is a no-op).
template <class L, int FORCE>
__device__ __noinline__ void neighbour_macro(const StepArgs &a, const int p, const int j, const int k, double &rho, double (&u)[3])
{
	const unsigned r = (unsigned)j * (unsigned)a.K + (unsigned)k;
	const long long id = (long long)p * a.MK + r;
	const uint32_t w = a.cw[id];
	const uint32_t cls = cw_class<L>(w);
	if (cls == CLS_FLUID || cls == CLS_GENERAL)      // upload guarantees the neighbour is eFluid or eSolid
	{
		double f[L::Q];
		pull_any<L>(a, p, r, j, k, id, w, f);
		macroscopic<L, FORCE>(f, a.hFg, rho, u);
	}
	else
	{
		rho = a.rho[id];
#pragma unroll
		for (int d = 0; d < 3; ++d) u[d] = (d < L::D) ? a.u[(long long)d * a.stride + id] : 0.0;
	}
}

// ------------------------------------------------------------------------------------------------
// velocity / pressure sites: stream, regularised boundary condition, (force), collide.
// GridObj::_LBM_regularised_opt (optimised.cpp:313-510).  bc_regularise() is everything between the
// pull and the store of one site; (r1,u1), (r2,u2) are the new-time moments of the two inward
// neighbours (only read for pressure faces and velocity edges/corners).
// ------------------------------------------------------------------------------------------------
template <class L>
__device__ __forceinline__ bool bc_needs_neighbours(const uint32_t w)
{
	return (w >> CW_EC_SHIFT) > 1u || cw_class<L>(w) == CLS_PRESSURE;
}

template <class L, int COLL, int FORCE>
__device__ __forceinline__ void bc_regularise(const StepArgs &a, const long long id, const uint32_t w, const int j, double (&f)[L::Q],
	const double r1, const double (&u1)[3], const double r2, const double (&u2)[3], double &dens, double (&uw)[3])
{
	const bool pressure = cw_class<L>(w) == CLS_PRESSURE;
	const int ec = (int)(w >> CW_EC_SHIFT);
	const int nd = (int)((w >> CW_ND_SHIFT) & 3u);
	int n[3];
#pragma unroll
	for (int d = 0; d < 3; ++d) n[d] = (int)((w >> (CW_N_SHIFT + 2 * d)) & 3u) - 1;
	const int nn = n[nd == 0 ? 0 : (nd == 1 ? 1 : 2)];

	// wall velocity (indexed by j whatever the wall, optimised.cpp:338-340) and reference density
	uw[0] = a.uin[j] * a.ramp; uw[1] = a.uin[a.M + j] * a.ramp; uw[2] = a.uin[2 * a.M + j] * a.ramp;
	dens = a.rho_out;

	if (ec > 1)
	{
		// edge / corner (velocity only): density extrapolated along the normal vector (:364)
		dens = 2.0 * r1 - r2;
	}
	else
	{
		double f_plus = 0.0, f_zero = 0.0;
#pragma unroll
		for (int v = 0; v < L::Q; ++v)
		{
			const int cn = (nd == 0) ? L::c(v, 0) : ((nd == 1) ? L::c(v, 1) : L::c(v, 2));
			if (cn == -nn) f_plus += f[v];
			else if (cn == 0) f_zero += f[v];
		}
		if (pressure)
		{
			// tangential velocity: first-order extrapolation of the new-time u (:396-402)
#pragma unroll
			for (int d = 0; d < L::D; ++d)
				if (d != nd) uw[d] = 2.0 * u1[d] - u2[d];
			double un = 1.0 - ((1.0 / dens) * (2.0 * f_plus + f_zero));
			if (nn == -1) un *= -1.0;
			if (nd == 0) uw[0] = un; else if (nd == 1) uw[1] = un; else uw[2] = un;
		}
		else
		{
			double un = (nd == 0) ? uw[0] : ((nd == 1) ? uw[1] : uw[2]);
			if (nn == -1) un *= -1.0;
			dens = (1.0 / (1.0 - un)) * (2.0 * f_plus + f_zero);
		}
	}

	double feq[L::Q];
	if (L::D == 2) uw[2] = 0.0;
	equilibrium_all<L>(dens, uw, a.C, feq);

	// non-equilibrium bounce-back on the unknown links, in ascending v with in-place updates (:435-489)
	double Sxx = 0.0, Syy = 0.0, Sxy = 0.0, Szz = 0.0, Sxz = 0.0, Syz = 0.0;
#pragma unroll
	for (int v = 0; v < L::Q; ++v)
	{
		const int c0 = L::c(v, 0), c1 = L::c(v, 1), c2 = L::c(v, 2);
		const int o = opposite<L>(v);
		const int cn = (nd == 0) ? c0 : ((nd == 1) ? c1 : c2);
		if (ec == 1)
		{
			if (cn == nn) f[v] = feq[v] + (f[o] - feq[o]);
		}
		else if (c0 == n[0] || c1 == n[1] || (L::D == 3 && c2 == n[2]))
		{
			const int dp = c0 * n[0] + c1 * n[1] + ((L::D == 3) ? c2 * n[2] : 0);
			const bool diagonal = (c0 * c0 + c1 * c1 + ((L::D == 3) ? c2 * c2 : 0)) > 1;   // sqrt(|c|^2) > 1.0
			if (dp == 0 && diagonal) f[v] = feq[v];                                      // buried link (:468-471)
			else f[v] = feq[v] + (f[o] - feq[o]);
		}
		const double fneq = f[v] - feq[v];
		Sxx += (double)(c0 * c0) * fneq;
		Syy += (double)(c1 * c1) * fneq;
		Sxy += (double)(c0 * c1) * fneq;
		if (L::D == 3)
		{
			Szz += (double)(c2 * c2) * fneq;
			Sxz += (double)(c0 * c2) * fneq;
			Syz += (double)(c1 * c2) * fneq;
		}
	}
	// regularised populations (:496-508)
#pragma unroll
	for (int v = 0; v < L::Q; ++v)
	{
		const int c0 = L::c(v, 0), c1 = L::c(v, 1), c2 = L::c(v, 2);
		f[v] = feq[v] + a.C.wden[L::wclass(v)] *
			(
			(((double)(c0 * c0) - a.C.cs2) * Sxx) +
			(((double)(c1 * c1) - a.C.cs2) * Syy) +
			(((double)(c2 * c2) - a.C.cs2) * Szz) +
			(2.0 * (double)c0 * (double)c1 * Sxy) +
			(2.0 * (double)c0 * (double)c2 * Sxz) +
			(2.0 * (double)c1 * (double)c2 * Syz)
			);
	}
	// macro is skipped for these types (:803-806); force and collide are applied (:138-150)
	collide<L, COLL, FORCE>(a, id, uw, feq, f);
}

// class-4 sites: stream per link, macro only for eFluid/eSlip (optimised.cpp:803-806; every other type
// keeps its stored rho,u), force, collide
template <class L, int COLL, int FORCE, bool TAVG>
__device__ __noinline__ void general_site(const StepArgs &a, const int p, const int j, const int k, const long long id,
	const uint32_t w, const int reps)
{
	const uint8_t type = a.types[id];
	double f[L::Q], feq[L::Q], u[3], rho;
	pull_general<L>(a, p, j, k, id, site_desc<L>(a, id, w), type, f);
	if (type == T_FLUID || type == T_SLIP) macroscopic<L, FORCE>(f, a.hFg, rho, u);
	else
	{
		rho = a.rho[id];
#pragma unroll
		for (int d = 0; d < 3; ++d) u[d] = (d < L::D) ? a.u[(long long)d * a.stride + id] : 0.0;
	}
	if (TAVG) tavg_update<L>(a, id, rho, u, reps);
	equilibrium_all<L>(rho, u, a.C, feq);
	collide<L, COLL, FORCE>(a, id, u, feq, f);
	store_populations<L>(a, id, f);
	store_outgoing<L>(a, p, (unsigned)j * (unsigned)a.K + (unsigned)k, f);
	a.rho[id] = rho;
#pragma unroll
	for (int d = 0; d < L::D; ++d) a.u[(long long)d * a.stride + id] = u[d];
}

template <class L, int COLL, int FORCE, bool TAVG>
__global__ void __launch_bounds__(64) k_bc(const StepArgs a)
{
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= a.n_bc) return;
	const long long id = a.bc_list[t];
	const int p = (int)(id / a.MK);
	const unsigned r = (unsigned)(id - (long long)p * a.MK);
	const int j = (int)(r / (unsigned)a.K), k = (int)(r - (unsigned)j * (unsigned)a.K);
	const uint32_t w = a.cw[id];
	const int reps = 1 + ((TAVG && a.bc_extra) ? a.bc_extra[t] : 0);
	if (!L::REGULARISABLE || cw_class<L>(w) == CLS_GENERAL)
	{
		general_site<L, COLL, FORCE, TAVG>(a, p, j, k, id, w, reps);
		return;
	}
	if constexpr (L::REGULARISABLE)
	{
	double f[L::Q];
	pull_any<L>(a, p, r, j, k, id, w, f);

	double r1 = 0.0, r2 = 0.0, u1[3] = { 0.0, 0.0, 0.0 }, u2[3] = { 0.0, 0.0, 0.0 };
	if (bc_needs_neighbours<L>(w))
	{
		int n[3];
#pragma unroll
		for (int d = 0; d < 3; ++d) n[d] = (int)((w >> (CW_N_SHIFT + 2 * d)) & 3u) - 1;
		neighbour_macro<L, FORCE>(a, p + n[0], j + n[1], k + n[2], r1, u1);
		neighbour_macro<L, FORCE>(a, p + 2 * n[0], j + 2 * n[1], k + 2 * n[2], r2, u2);
	}
	double dens, uw[3];
	bc_regularise<L, COLL, FORCE>(a, id, w, j, f, r1, u1, r2, u2, dens, uw);
	if (TAVG) tavg_update<L>(a, id, dens, uw, reps);      // _LBM_macro_opt still runs its averaging tail for these types

	a.rho[id] = dens;
#pragma unroll
	for (int d = 0; d < L::D; ++d) a.u[(long long)d * a.stride + id] = uw[d];
	store_populations<L>(a, id, f);
	store_outgoing<L>(a, p, r, f);
	}
}

// ------------------------------------------------------------------------------------------------
// Forced-equilibrium velocity BC with L_VELOCITY_RAMP: every site that pulls from an eVelocity
// site first overwrites that site's stored u with u_in[j of the PULLING site] * ramp (optimised.cpp:
// 258-263), so what the array holds after a step is the value written by the LAST puller in the
// reference's loop order (x slowest, then y, then z).  The populations never depend on the stored
// value (every puller uses its own), so this kernel runs once, after the last step of a call.
// ------------------------------------------------------------------------------------------------
template <class L>
__global__ void __launch_bounds__(64) k_velsrc(const VelSrcArgs a)
{
	const int t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= a.n) return;
	const long long MK = (long long)a.M * a.K;
	const long long id = a.list[t];
	const int p = (int)(id / MK);
	const int r = (int)(id - (long long)p * MK);
	const int j = r / a.K, k = r - j * a.K;
	long long best = -1;
	int best_j = 0;
#pragma unroll
	for (int v = 0; v < L::Q - 1; ++v)
	{
		const int cx = L::c(v, 0), cy = L::c(v, 1), cz = L::c(v, 2);
		int dp = p + cx, dj = j + cy, dk = k + cz;
		if (a.wrap_x) { if (dp < 0) dp += a.P; else if (dp >= a.P) dp -= a.P; }
		else if (dp < 0 || dp >= a.P) continue;
		if (dj < 0) dj += a.M; else if (dj >= a.M) dj -= a.M;
		if (dk < 0) dk += a.K; else if (dk >= a.K) dk -= a.K;
		const long long did = ((long long)dp * a.M + dj) * a.K + dk;
		const uint8_t dt = a.types[did];
		if (dt == T_SOLID || dt == T_REFINED || dt == T_VELOCITY) continue;      // never streamed (optimised.cpp:91-95)
		if (dt == T_SLIP)
		{
			// a slip site reflects this link instead of pulling it (:229-233)
			const uint32_t d = a.bcdesc[did];
			const int n0 = (int)((d >> CW_N_SHIFT) & 3u) - 1, n1 = (int)((d >> (CW_N_SHIFT + 2)) & 3u) - 1, n2 = (int)((d >> (CW_N_SHIFT + 4)) & 3u) - 1;
			if ((cx != 0 && n0 == cx) || (cy != 0 && n1 == cy) || (L::D == 3 && cz != 0 && n2 == cz)) continue;
		}
		int gx = (a.x_first + dp) % a.N;
		if (gx < 0) gx += a.N;
		const long long key = ((long long)gx * a.M + dj) * a.K + dk;
		if (key > best) { best = key; best_j = dj; }
	}
	if (best >= 0)
	{
#pragma unroll
		for (int d = 0; d < L::D; ++d) a.u[(long long)d * a.stride + id] = a.uin[d * a.M + best_j] * a.ramp_t;
	}
}

#ifdef __CUDACC__   // launchers: <<<>>> needs nvcc (tests/harness compiles the kernels above for the host)
template <class L> void launch_velsrc(const VelSrcArgs &a, cudaStream_t s, int64_t *launches)
{
	if (a.n <= 0) return;
	k_velsrc<L><<<(a.n + 63) / 64, 64, 0, s>>>(a);
	if (launches) ++*launches;
}
#endif

#ifdef __CUDACC__   // launchers: <<<>>> needs nvcc (tests/harness compiles the kernels above for the host)
// kernel variant = collision operator x Guo forcing x time averages; the reference ties KBC to D2Q9 / D3Q27 and
// D3Q27 to KBC (inc/definitions.h:299-310), so only those combinations are instantiated
#define LUMA_DISPATCH_T(KERNEL, COLL, FORCE_, GRID, THREADS) \
	do { \
		if (a.tav) KERNEL<L, COLL, FORCE_, true><<<GRID, THREADS, 0, s>>>(a); \
		else KERNEL<L, COLL, FORCE_, false><<<GRID, THREADS, 0, s>>>(a); \
	} while (0)
// force: 0 = none, 1 + L_GRAVITY_DIRECTION otherwise (the direction is a template argument: the zero components of
// force_xyz vanish at compile time)
#define LUMA_DISPATCH_FT(KERNEL, COLL, GRID, THREADS) \
	do { \
		switch (force) \
		{ \
		case 0: LUMA_DISPATCH_T(KERNEL, COLL, 0, GRID, THREADS); break; \
		case 1: LUMA_DISPATCH_T(KERNEL, COLL, 1, GRID, THREADS); break; \
		case 2: LUMA_DISPATCH_T(KERNEL, COLL, 2, GRID, THREADS); break; \
		default: if constexpr (L::D == 3) { LUMA_DISPATCH_T(KERNEL, COLL, 3, GRID, THREADS); } break; \
		} \
	} while (0)
#define LUMA_DISPATCH(KERNEL, GRID, THREADS) \
	do { \
		if constexpr (L::Q == 27) { LUMA_DISPATCH_FT(KERNEL, COLL_KBC, GRID, THREADS); } \
		else if (coll == COLL_KBC) { if constexpr (L::Q == 9) { LUMA_DISPATCH_FT(KERNEL, COLL_KBC, GRID, THREADS); } } \
		else if (coll == COLL_SMAG) { LUMA_DISPATCH_FT(KERNEL, COLL_SMAG, GRID, THREADS); } \
		else { LUMA_DISPATCH_FT(KERNEL, COLL_BGK, GRID, THREADS); } \
	} while (0)

template <class L> void launch_step(const StepArgs &a, int coll, int force, int nplanes, cudaStream_t s, int64_t *launches)
{
	if (nplanes <= 0) return;
	dim3 grid((a.MK + STEP_THREADS - 1) / STEP_THREADS, (unsigned)nplanes);
	// the TMA-staged variant needs even M*K and K (16-byte aligned bulk copies) and the start of a tile on an even element
#if LUMA_VARIANTS
	if (a.use_tma == 1 && (a.MK & 1u) == 0 && (a.K & 1) == 0) LUMA_DISPATCH(k_step_tma, grid, STEP_THREADS);
	else if (a.use_tma == 2 && (a.MK & 1u) == 0 && (a.K & 1) == 0 && L::D == 3)
	{
		// two sites per thread: half as many threads per plane
		dim3 grid2((a.MK / 2 + STEP_THREADS - 1) / STEP_THREADS, (unsigned)nplanes);
		LUMA_DISPATCH(k_step_v2, grid2, STEP_THREADS);
		StepArgs rest = a;
		rest.rest_only = 1;
		{
			const StepArgs &a = rest;
			LUMA_DISPATCH(k_step, grid, STEP_THREADS);
		}
		if (launches) ++*launches;
	}
	else
#endif
	LUMA_DISPATCH(k_step, grid, STEP_THREADS);
	if (launches) ++*launches;
}

template <class L> void launch_step_faces(const StepArgs &a, int coll, int force, int nplanes, cudaStream_t s, int64_t *launches)
{
	if (nplanes <= 0) return;
	dim3 grid((a.MK + STEP_THREADS - 1) / STEP_THREADS, (unsigned)nplanes);
	LUMA_DISPATCH(k_step_faces, grid, STEP_THREADS);
	if (launches) ++*launches;
}

template <class L> void launch_bc(const StepArgs &a, int coll, int force, cudaStream_t s, int64_t *launches)
{
	if (a.n_bc <= 0) return;
	const int threads = 64;
	dim3 grid((a.n_bc + threads - 1) / threads);
	LUMA_DISPATCH(k_bc, grid, threads);
	if (launches) ++*launches;
}
#endif

// ------------------------------------------------------------------------------------------------
// geometry: cell words from the eType array (+ host- or device-made wall descriptors)
// ------------------------------------------------------------------------------------------------
template <class L>
__global__ void k_cell_words(const GeomArgs g)
{
	const unsigned MK = (unsigned)g.M * (unsigned)g.K;
	const unsigned r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= MK) return;
	const int p = g.p_begin + (int)blockIdx.y;
	const int j = (int)(r / (unsigned)g.K), k = (int)(r - (unsigned)j * (unsigned)g.K);
	const long long id = (long long)p * MK + r;
	const uint8_t t = g.types[id];
	uint32_t cls = CLS_SKIP;
	if (t == T_FLUID) cls = CLS_FLUID;
	else if (t == T_VELOCITY) cls = g.regularised ? CLS_VELOCITY : CLS_SKIP;      // optimised.cpp:91-95
	else if (t == T_PRESSURE) cls = g.regularised ? CLS_PRESSURE : CLS_GENERAL;
	else if (t == T_SLIP || t == T_EXTRAPOLATE_RIGHT) cls = CLS_GENERAL;
	uint32_t w = 0;
	if (cls != CLS_SKIP)
	{
		w = cls << CW<L>::CLASS_SHIFT;
#pragma unroll
		for (int v = 0; v < L::Q - 1; ++v)
		{
			int sp = p - L::c(v, 0), sj = j - L::c(v, 1), sk = k - L::c(v, 2);
			if (g.wrap_x) sp = (sp + g.P) % g.P;
			sj = (sj + g.M) % g.M;
			sk = (sk + g.K) % g.K;
			const long long src = ((long long)sp * g.M + sj) * g.K + sk;
			if (g.types[src] == 0) w |= 1u << v;
		}
		if (j == 0 || j == g.M - 1 || (L::D == 3 && (k == 0 || k == g.K - 1))) w |= CW<L>::EDGE;
		if (L::DESC_IN_WORD && cls >= CLS_VELOCITY && g.bcdesc) w |= g.bcdesc[id];
	}
	g.cw[id] = w;
}

#ifdef __CUDACC__   // launchers: <<<>>> needs nvcc (tests/harness compiles the kernels above for the host)
template <class L> void launch_cell_words(const GeomArgs &g, cudaStream_t s)
{
	const unsigned MK = (unsigned)g.M * (unsigned)g.K;
	dim3 grid((MK + 255) / 256, (unsigned)(g.p_end - g.p_begin));
	if (g.p_end > g.p_begin) k_cell_words<L><<<grid, 256, 0, s>>>(g);
}
#endif

// ------------------------------------------------------------------------------------------------
// device-side LBM_initGrid for index-described cases (src/GridObj_init_grids.cpp:155-384):
// labels (LBM_initBoundLab :983-1097, precedence :1372-1377), u/rho (:37-150), f = feq (:310-333),
// then the bounce-back body (src/ObjectManager.cpp:309-345), and the wall descriptors
// (GridUtils::isWithinDomainWall, src/GridUtils.cpp:1369-1430) in cell-index form.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int bc_precedence(int current, int desired)
{
	if (current == 0 || desired == 0) return 0;
	if (current == 6) return 6;
	return desired;
}

template <class L>
__global__ void k_synthetic(const SynthArgs a)
{
	const unsigned MK = (unsigned)a.M * (unsigned)a.K;
	const unsigned r = blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= MK) return;
	const int p = (int)blockIdx.y;
	const int j = (int)(r / (unsigned)a.K), k = (int)(r - (unsigned)j * (unsigned)a.K);
	const long long id = (long long)p * MK + r;
	const int gi = ((a.x_first + p) % a.N + a.N) % a.N;
	const int *wc = a.wall_cells, *wt = a.wall_type;

	int type = 1;
	if (gi < wc[0]) type = bc_precedence(type, wt[0]);
	if (gi >= a.N - wc[1]) type = bc_precedence(type, wt[1]);
	if (L::D == 3)
	{
		if (k < wc[4]) type = bc_precedence(type, wt[4]);
		if (k >= a.K - wc[5]) type = bc_precedence(type, wt[5]);
	}
	if (j < wc[2]) type = bc_precedence(type, wt[2]);
	if (j >= a.M - wc[3]) type = bc_precedence(type, wt[3]);

	double u[3] = { 0.0, 0.0, 0.0 };
	if (!(a.no_flow && type != 6))
	{
#pragma unroll
		for (int d = 0; d < L::D; ++d) u[d] = a.uin[d * a.M + j] * a.ramp0;
	}
	if (type == 0) { u[0] = 0.0; u[1] = 0.0; u[2] = 0.0; }
	const double rho = a.rhoin;

	double feq[L::Q];
	equilibrium_all<L>(rho, u, a.C, feq);
#pragma unroll
	for (int v = 0; v < L::Q; ++v)
	{
		a.f0[(long long)v * a.stride + id] = feq[v];
		a.f1[(long long)v * a.stride + id] = feq[v];
	}

	if (a.has_box && type == 1 && gi >= a.box[0] && gi < a.box[1] && j >= a.box[2] && j < a.box[3] && k >= a.box[4] && k < a.box[5])
	{
		type = 0;
		u[0] = 0.0; u[1] = 0.0;   // the reference never zeroes the z component (ObjectManager.cpp:333-338)
	}
	a.types[id] = (uint8_t)type;
	a.rho[id] = rho;
#pragma unroll
	for (int d = 0; d < L::D; ++d) a.u[(long long)d * a.stride + id] = u[d];

	int ec = 0, nd = 0, n0 = 0, n1 = 0, n2 = 0;
	if (gi < wc[0]) { nd = 0; n0 = 1; ++ec; }
	if (gi >= a.N - wc[1]) { nd = 0; n0 = -1; ++ec; }
	if (j < wc[2]) { nd = 1; n1 = 1; ++ec; }
	if (j >= a.M - wc[3]) { nd = 1; n1 = -1; ++ec; }
	if (L::D == 3)
	{
		if (k < wc[4]) { nd = 2; n2 = 1; ++ec; }
		if (k >= a.K - wc[5]) { nd = 2; n2 = -1; ++ec; }
	}
	a.bcdesc[id] = (ec > 0) ? cw_pack_bc(ec, nd, n0, n1, n2) : 0u;
}

#ifdef __CUDACC__   // launchers: <<<>>> needs nvcc (tests/harness compiles the kernels above for the host)
template <class L> void launch_synthetic(const SynthArgs &a, cudaStream_t s)
{
	const unsigned MK = (unsigned)a.M * (unsigned)a.K;
	dim3 grid((MK + 127) / 128, (unsigned)a.P);
	k_synthetic<L><<<grid, 128, 0, s>>>(a);
}
#endif

// ------------------------------------------------------------------------------------------------
// f = feq(rho, u) at every site of a plane range: the population initialisation of LBM_initGrid
// (src/GridObj_init_grids.cpp:310-333) on the device, for hosts that hand over rho, u and LatTyp of a
// freshly initialised grid (t = 0) and keep the 19 x 8 B per site of f off the PCIe bus.
// ------------------------------------------------------------------------------------------------
template <class L>
__global__ void __launch_bounds__(128) k_feq_init(const double *__restrict__ rho, const double *__restrict__ u, double *__restrict__ f,
	const long long stride, const long long first, const long long n, const LbmConst C)
{
	const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= n) return;
	const long long id = first + s;
	double uu[3] = { 0.0, 0.0, 0.0 }, feq[L::Q];
#pragma unroll
	for (int d = 0; d < L::D; ++d) uu[d] = u[(long long)d * stride + id];
	equilibrium_all<L>(rho[id], uu, C, feq);
#pragma unroll
	for (int v = 0; v < L::Q; ++v) f[(long long)v * stride + id] = feq[v];
}

#ifdef __CUDACC__
template <class L> void launch_feq_init(const double *rho, const double *u, double *f, long long stride, long long first, long long n,
	const LbmConst &C, cudaStream_t s)
{
	if (n > 0) k_feq_init<L><<<(unsigned)((n + 127) / 128), 128, 0, s>>>(rho, u, f, stride, first, n, C);
}
#endif

// ------------------------------------------------------------------------------------------------
// layout conversion between LUMA's AoS (inc/IVector.h:94-134) and the device SoA
// ------------------------------------------------------------------------------------------------
template <int Q>
__global__ void k_aos_to_soa(const double *__restrict__ aos, double *__restrict__ soa, long long stride, long long first, long long n)
{
	__shared__ double tile[Q * 64];
	const long long c0 = (long long)blockIdx.x * 64;
	const int cnt = (int)((n - c0) < 64 ? (n - c0) : 64);
	for (int e = threadIdx.x; e < cnt * Q; e += blockDim.x) tile[e] = aos[c0 * Q + e];
	__syncthreads();
	for (int e = threadIdx.x; e < cnt * Q; e += blockDim.x)
	{
		const int v = e / cnt, c = e - v * cnt;
		soa[(long long)v * stride + first + c0 + c] = tile[c * Q + v];
	}
}

template <int Q>
__global__ void k_soa_to_aos(const double *__restrict__ soa, double *__restrict__ aos, long long stride, long long first, long long n)
{
	__shared__ double tile[Q * 64];
	const long long c0 = (long long)blockIdx.x * 64;
	const int cnt = (int)((n - c0) < 64 ? (n - c0) : 64);
	for (int e = threadIdx.x; e < cnt * Q; e += blockDim.x)
	{
		const int v = e / cnt, c = e - v * cnt;
		tile[c * Q + v] = soa[(long long)v * stride + first + c0 + c];
	}
	__syncthreads();
	for (int e = threadIdx.x; e < cnt * Q; e += blockDim.x) aos[c0 * Q + e] = tile[e];
}

#ifdef __CUDACC__   // launchers: <<<>>> needs nvcc (tests/harness compiles the kernels above for the host)
template <class L> void launch_aos_to_soa(const double *aos, double *soa, long long stride, long long first, long long n, cudaStream_t s)
{
	if (n > 0) k_aos_to_soa<L::Q><<<(unsigned)((n + 63) / 64), 256, 0, s>>>(aos, soa, stride, first, n);
}
template <class L> void launch_soa_to_aos(const double *soa, double *aos, long long stride, long long first, long long n, cudaStream_t s)
{
	if (n > 0) k_soa_to_aos<L::Q><<<(unsigned)((n + 63) / 64), 256, 0, s>>>(soa, aos, stride, first, n);
}
#endif

// ------------------------------------------------------------------------------------------------
// momentum exchange on eSolid sites, ObjectManager::computeLiftDrag(i,j,k,g)
// (src/ObjectManager.cpp:93-164): for every link n of a solid site whose far end (site - c_opp(n))
// is on the grid and eFluid, add 2 c_opp f_opp(far end), f = populations BEFORE the step.
// Slabs: a link is summed by the rank that OWNS ITS FLUID END -- the solid end may lie in a ghost plane (eType of
// the ghost planes is exchanged once, with the geometry), the populations read are always those of owned planes.
// No ghost-plane population is read, so a ring neighbour that already runs the next step (and stores into the
// ghost planes of the lattice read here) cannot race with this kernel.
// Summation order here: per-thread over n ascending, fixed-shape tree over the block, then the
// host adds the per-block partials in block order (deterministic; differs from the reference's
// serial i,j,k order, hence a tolerance in the tests).
// ------------------------------------------------------------------------------------------------
template <class L>
__global__ void __launch_bounds__(256) k_momex(const double *__restrict__ f, const uint8_t *__restrict__ types, long long stride,
	int P, int M, int K, int p_begin, int p_end, int x_first, int N, double *__restrict__ partials)
{
	const long long MK = (long long)M * K;
	const long long total = (long long)P * MK;
	double F0 = 0.0, F1 = 0.0, F2 = 0.0;
	for (long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x; s < total; s += (long long)gridDim.x * blockDim.x)
	{
		const int p = (int)(s / MK);
		const long long r = s - (long long)p * MK;
		const int j = (int)(r / K), k = (int)(r - (long long)j * K);
		const int gi = x_first + p;
		if (gi < 0 || gi >= N) continue;      // ghost plane beyond the grid (the periodic image: no momentum exchange across the wrap)
		if (types[s] != 0) continue;
#pragma unroll
		for (int n = 0; n < L::Q; ++n)
		{
			const int no = opposite<L>(n);
			const int xd = gi - L::c(no, 0), yd = j - L::c(no, 1), zd = k - L::c(no, 2);
			if (xd < 0 || xd >= N || yd < 0 || yd >= M || zd < 0 || zd >= K) continue;
			const int dp = p - L::c(no, 0);
			if (dp < p_begin || dp >= p_end) continue;      // the fluid end belongs to another rank
			const long long dest = ((long long)dp * M + yd) * K + zd;
			if (types[dest] != 1) continue;
			const double fv = f[(long long)no * stride + dest];
			F0 += 2.0 * (double)L::c(no, 0) * fv;
			F1 += 2.0 * (double)L::c(no, 1) * fv;
			F2 += 2.0 * (double)L::c(no, 2) * fv;
		}
	}
	__shared__ double sh[3][256];
	sh[0][threadIdx.x] = F0; sh[1][threadIdx.x] = F1; sh[2][threadIdx.x] = F2;
	__syncthreads();
	for (int w = 128; w > 0; w >>= 1)
	{
		if ((int)threadIdx.x < w)
		{
			sh[0][threadIdx.x] += sh[0][threadIdx.x + w];
			sh[1][threadIdx.x] += sh[1][threadIdx.x + w];
			sh[2][threadIdx.x] += sh[2][threadIdx.x + w];
		}
		__syncthreads();
	}
	if (threadIdx.x == 0)
	{
		partials[3 * blockIdx.x + 0] = sh[0][0];
		partials[3 * blockIdx.x + 1] = sh[1][0];
		partials[3 * blockIdx.x + 2] = sh[2][0];
	}
}

#ifdef __CUDACC__   // launchers: <<<>>> needs nvcc (tests/harness compiles the kernels above for the host)
template <class L> int launch_momex(const double *f_prev, const uint8_t *types, long long stride, int P, int M, int K,
	int p_begin, int p_end, int x_first, int N, double *partials, int max_blocks, cudaStream_t s)
{
	const long long total = (long long)P * M * K;
	int blocks = (int)((total + 255) / 256);
	if (blocks > max_blocks) blocks = max_blocks;
	if (blocks < 1) blocks = 1;
	k_momex<L><<<blocks, 256, 0, s>>>(f_prev, types, stride, P, M, K, p_begin, p_end, x_first, N, partials);
	return blocks;
}
#endif

// explicit instantiations
#define LUMA_INST(L) \
	template void launch_step<L>(const StepArgs &, int, int, int, cudaStream_t, int64_t *); \
	template void launch_bc<L>(const StepArgs &, int, int, cudaStream_t, int64_t *); \
	template void launch_step_faces<L>(const StepArgs &, int, int, int, cudaStream_t, int64_t *); \
	template void launch_velsrc<L>(const VelSrcArgs &, cudaStream_t, int64_t *); \
	template void launch_cell_words<L>(const GeomArgs &, cudaStream_t); \
	template void launch_synthetic<L>(const SynthArgs &, cudaStream_t); \
	template void launch_feq_init<L>(const double *, const double *, double *, long long, long long, long long, const LbmConst &, cudaStream_t); \
	template void launch_aos_to_soa<L>(const double *, double *, long long, long long, long long, cudaStream_t); \
	template void launch_soa_to_aos<L>(const double *, double *, long long, long long, long long, cudaStream_t); \
	template int launch_momex<L>(const double *, const uint8_t *, long long, int, int, int, int, int, int, int, double *, int, cudaStream_t);

}  // namespace luma
