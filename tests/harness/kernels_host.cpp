// TEST INFRASTRUCTURE ONLY -- compiles the product's kernel bodies (luma_b200/csrc/kernels_impl.cuh: k_cell_words,
// k_step, k_bc, k_velsrc with everything they call) for the HOST and runs them thread by thread in plain loops, so
// that tests/test_kernels_host_emulation.py can replay every parity case against the oracle without a GPU.  It is a
// logic check of the device code (index arithmetic, branch ladders, operation order), not an implementation anyone can
// use: one "thread" at a time, single slab, no streams, no exchange, no C ABI.  Nothing in the product links or loads
// this file, and the product has no CPU path.  Built with g++ -O2 -ffp-contract=off (no contraction, like -fmad=false).
#include <cmath>
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>      // host-side declarations only (dim3, cudaStream_t)

// what nvcc provides inside kernels
struct EmuIdx { unsigned x, y, z; };
static thread_local EmuIdx blockIdx, threadIdx, blockDim, gridDim;
template <class T> static inline T __ldg(const T *p) { return *p; }
static inline void __syncthreads() {}
#undef __noinline__
#define __noinline__
#undef __launch_bounds__
#define __launch_bounds__(...)

#include "../../luma_b200/csrc/kernels_impl.cuh"

using namespace luma;

// same expressions as make_constants() in luma_b200/csrc/api.cu
static void make_constants(LbmConst &C, int Q)
{
	const volatile double three = 3.0, one = 1.0;
	const double cs = one / std::sqrt(three);
	C.cs2 = cs * cs;
	C.inv_cs2 = 1.0 / C.cs2;
	C.den = (2.0 * C.cs2) * C.cs2;
	C.inv_den = 1.0 / C.den;
	C.k1 = 1.0 - C.cs2;
	C.k0 = 0.0 - C.cs2;
	C.w[3] = 0.0;
	if (Q == 27) { C.w[0] = 2.0 / 27.0; C.w[1] = 1.0 / 54.0; C.w[2] = 1.0 / 216.0; C.w[3] = 8.0 / 27.0; }
	else if (Q == 19) { C.w[0] = 1.0 / 18.0; C.w[1] = 1.0 / 36.0; C.w[2] = 1.0 / 3.0; }
	else { C.w[0] = 1.0 / 9.0; C.w[1] = 1.0 / 36.0; C.w[2] = 4.0 / 9.0; }
	for (int k = 0; k < 4; ++k) C.wden[k] = C.w[k] / C.den;
}

template <class Fn>
static void emu_launch(unsigned gx, unsigned gy, unsigned threads, Fn fn)
{
	gridDim.x = gx; gridDim.y = gy; gridDim.z = 1;
	blockDim.x = threads; blockDim.y = 1; blockDim.z = 1;
	for (unsigned by = 0; by < gy; ++by)
		for (unsigned bx = 0; bx < gx; ++bx)
			for (unsigned t = 0; t < threads; ++t)
			{
				blockIdx.x = bx; blockIdx.y = by; blockIdx.z = 0;
				threadIdx.x = t; threadIdx.y = 0; threadIdx.z = 0;
				fn();
			}
}

template <class L, int COLL, int FORCE, bool TAVG>
static void step_kernels(const StepArgs &a, unsigned gx, unsigned gy, unsigned gbc, bool run_bc, bool faces)
{
	// k_bc first, k_step second: they read fin and write disjoint sites of fout (any order gives the same result)
	if (a.n_bc > 0 && run_bc) emu_launch(gbc, 1, 64, [&] { k_bc<L, COLL, FORCE, TAVG>(a); });
	if (gy == 0) return;
	if (faces) emu_launch(gx, gy, STEP_THREADS, [&] { k_step_faces<L, COLL, FORCE, TAVG>(a); });
	else emu_launch(gx, gy, STEP_THREADS, [&] { k_step<L, COLL, FORCE, TAVG>(a); });
}

// force: 0 none, 1 + L_GRAVITY_DIRECTION (the dispatch of LUMA_DISPATCH_FT in kernels_impl.cuh)
template <class L, int COLL>
static void step_ft(const StepArgs &a, int force, unsigned gx, unsigned gy, unsigned gbc, bool run_bc, bool faces)
{
#define EMU_T(F_) do { if (a.tav) step_kernels<L, COLL, F_, true>(a, gx, gy, gbc, run_bc, faces); else step_kernels<L, COLL, F_, false>(a, gx, gy, gbc, run_bc, faces); } while (0)
	switch (force)
	{
	case 0: EMU_T(0); break;
	case 1: EMU_T(1); break;
	case 2: EMU_T(2); break;
	default: if constexpr (L::D == 3) { EMU_T(3); } break;
	}
#undef EMU_T
}

extern "C" {

struct EmuCase
{
	int32_t Q, D, P, M, K;          // P = local planes (with the two ghost planes when wrap_x == 0)
	int32_t regularised, coll, force, gravity_dir, velramp_on, general;
	double omega, rhoin, rho_out, gravity, csmag;
	double ramp, ramp_t, t_now, t_next;
	int32_t wrap_x;                 // 1: single slab, periodic wrap inside the array; 0: ghost planes 0 and P-1
	int32_t p0, pstep, nplanes;     // planes this emu_step call covers with k_step: p0 + n * pstep, n < nplanes
	int32_t run_bc;                 // run k_bc in this call
	int32_t N, x_first;             // global x size and global x of local plane 0 (k_velsrc, k_synthetic)
	// fused halo exchange: the two neighbour slabs' output lattices (here: other arrays of the same process)
	int32_t faces;                  // run k_step_faces instead of k_step
	int32_t peer_P[2];
	long long peer_stride[2];
	double *peer_f[2];
};

int emu_class_shift(int Q) { return Q == 27 ? CW<D3Q27>::CLASS_SHIFT : CW<D3Q19>::CLASS_SHIFT; }

int emu_cell_words(const EmuCase *c, const uint8_t *types, const uint32_t *bcdesc, uint32_t *cw)
{
	GeomArgs g;
	memset(&g, 0, sizeof(g));
	g.types = types; g.bcdesc = bcdesc; g.cw = cw;
	const int ghost = c->wrap_x ? 0 : 1;
	g.P = c->P; g.M = c->M; g.K = c->K; g.wrap_x = c->wrap_x; g.p_begin = ghost; g.p_end = c->P - ghost; g.regularised = c->regularised;
	const unsigned MK = (unsigned)c->M * (unsigned)c->K;
	const unsigned gx = (MK + 255) / 256, gy = (unsigned)(g.p_end - g.p_begin);
	if (c->Q == 19) emu_launch(gx, gy, 256, [&] { k_cell_words<D3Q19>(g); });
	else if (c->Q == 27) emu_launch(gx, gy, 256, [&] { k_cell_words<D3Q27>(g); });
	else emu_launch(gx, gy, 256, [&] { k_cell_words<D2Q9>(g); });
	return 0;
}

// one time step of a single slab (the argument block is filled the way luma_b200_step fills it, api.cu)
int emu_step(const EmuCase *c, const double *fin, double *fout, const uint32_t *cw, double *rho, double *u, long long stride,
	const long long *bc_list, const int *bc_extra, int n_bc, const double *uin, const uint8_t *types, const uint32_t *bcdesc, double *tav)
{
	StepArgs a;
	memset(&a, 0, sizeof(a));
	make_constants(a.C, c->Q);
	a.fin = fin; a.fout = fout; a.cw = cw; a.rho = rho; a.u = u; a.stride = stride;
	a.P = c->P; a.M = c->M; a.K = c->K; a.MK = (unsigned)c->M * (unsigned)c->K;
	a.wrap_x = c->wrap_x; a.p0 = c->p0; a.pstep = c->pstep; a.write_macro = 1;
	for (int v = 0; v < c->Q; ++v)
	{
		const int cx = c->Q == 19 ? D3Q19::c(v, 0) : (c->Q == 27 ? D3Q27::c(v, 0) : D2Q9::c(v, 0));
		const int cy = c->Q == 19 ? D3Q19::c(v, 1) : (c->Q == 27 ? D3Q27::c(v, 1) : D2Q9::c(v, 1));
		const int cz = c->Q == 19 ? D3Q19::c(v, 2) : (c->Q == 27 ? D3Q27::c(v, 2) : D2Q9::c(v, 2));
		a.off_pull[v] = 8LL * ((long long)v * stride - ((long long)cx * a.MK + (long long)cy * c->K + cz));
	}
	a.bc_list = bc_list; a.bc_extra = bc_extra; a.n_bc = n_bc; a.uin = uin;
	a.rho_out = c->rho_out;
	a.types = types; a.bcdesc = bcdesc; a.general = c->general; a.regularised = c->regularised; a.velramp_on = c->velramp_on;
	a.tav = tav;
	if (c->force) { a.Fg = c->rhoin * c->gravity * 1.0; a.hFg = 0.5 * a.Fg; }
	a.smag_coef = 2.0 * 1.4142135623730950488016887242097 * (c->csmag * c->csmag) * c->rhoin * a.C.cs2 * a.C.cs2;
	a.omega = c->omega;
	a.tau = 1.0 / c->omega;
	for (int k = 0; k < 4; ++k) a.lam[k] = (1 - 0.5 * c->omega) * (a.C.w[k] / a.C.cs2);
	a.kbc_beta_m1 = 2.0 / c->omega;
	a.kbc_inv_beta = 1.0 / a.kbc_beta_m1;
	a.ramp = c->ramp; a.ramp_t = c->ramp_t; a.t_now = c->t_now; a.t_next = c->t_next;
	for (int side = 0; side < 2; ++side) { a.peer_f[side] = c->peer_f[side]; a.peer_stride[side] = c->peer_stride[side]; a.peer_P[side] = c->peer_P[side]; }

	const unsigned gx = (a.MK + STEP_THREADS - 1) / STEP_THREADS, gy = (unsigned)(c->nplanes > 0 ? c->nplanes : 0), gbc = (unsigned)((n_bc + 63) / 64);
	const int force = c->force ? 1 + c->gravity_dir : 0;
	const bool bc = c->run_bc != 0, fc = c->faces != 0;
	if (c->Q == 27) step_ft<D3Q27, COLL_KBC>(a, force, gx, gy, gbc, bc, fc);
	else if (c->Q == 19)
	{
		if (c->coll == COLL_KBC) return 1;
		if (c->coll == COLL_SMAG) step_ft<D3Q19, COLL_SMAG>(a, force, gx, gy, gbc, bc, fc);
		else step_ft<D3Q19, COLL_BGK>(a, force, gx, gy, gbc, bc, fc);
	}
	else
	{
		if (c->coll == COLL_KBC) step_ft<D2Q9, COLL_KBC>(a, force, gx, gy, gbc, bc, fc);
		else if (c->coll == COLL_SMAG) step_ft<D2Q9, COLL_SMAG>(a, force, gx, gy, gbc, bc, fc);
		else step_ft<D2Q9, COLL_BGK>(a, force, gx, gy, gbc, bc, fc);
	}
	return 0;
}

// stored u of the forced-equilibrium inlet sites after a step (k_velsrc)
int emu_velsrc(const EmuCase *c, const long long *list, int n, const uint8_t *types, const uint32_t *bcdesc, double *u, long long stride,
	const double *uin)
{
	if (n <= 0) return 0;
	VelSrcArgs a;
	memset(&a, 0, sizeof(a));
	a.list = list; a.n = n; a.types = types; a.bcdesc = bcdesc; a.u = u; a.stride = stride; a.uin = uin; a.ramp_t = c->ramp_t;
	a.P = c->P; a.M = c->M; a.K = c->K; a.N = c->N; a.wrap_x = c->wrap_x; a.x_first = c->x_first;
	const unsigned g = (unsigned)((n + 63) / 64);
	if (c->Q == 19) emu_launch(g, 1, 64, [&] { k_velsrc<D3Q19>(a); });
	else if (c->Q == 27) emu_launch(g, 1, 64, [&] { k_velsrc<D3Q27>(a); });
	else emu_launch(g, 1, 64, [&] { k_velsrc<D2Q9>(a); });
	return 0;
}

// device-side LBM_initGrid (k_synthetic) of one slab
int emu_synthetic(const EmuCase *c, const int32_t *wall_type, const int32_t *wall_cells, const double *uin, double ramp0, int no_flow,
	int has_box, const int32_t *box, uint8_t *types, uint32_t *bcdesc, double *f0, double *f1, double *rho, double *u, long long stride)
{
	SynthArgs a;
	memset(&a, 0, sizeof(a));
	a.types = types; a.bcdesc = bcdesc; a.f0 = f0; a.f1 = f1; a.rho = rho; a.u = u; a.stride = stride;
	a.P = c->P; a.M = c->M; a.K = c->K; a.N = c->N; a.x_first = c->x_first;
	for (int i = 0; i < 6; ++i) { a.wall_type[i] = wall_type[i]; a.wall_cells[i] = wall_cells[i]; a.box[i] = box ? box[i] : 0; }
	a.uin = uin; a.ramp0 = ramp0; a.rhoin = c->rhoin; a.no_flow = no_flow; a.has_box = has_box;
	make_constants(a.C, c->Q);
	const unsigned MK = (unsigned)c->M * (unsigned)c->K;
	const unsigned gx = (MK + 127) / 128;
	if (c->Q == 19) emu_launch(gx, c->P, 128, [&] { k_synthetic<D3Q19>(a); });
	else if (c->Q == 27) emu_launch(gx, c->P, 128, [&] { k_synthetic<D3Q27>(a); });
	else emu_launch(gx, c->P, 128, [&] { k_synthetic<D2Q9>(a); });
	return 0;
}

int emu_lattice_c(int Q, int v, int d) { return Q == 9 ? D2Q9::c(v, d) : (Q == 19 ? D3Q19::c(v, d) : D3Q27::c(v, d)); }

}  // extern "C"
