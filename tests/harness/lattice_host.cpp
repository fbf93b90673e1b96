// TEST INFRASTRUCTURE ONLY -- compiles the product's arithmetic header (luma_b200/csrc/lattice.cuh) for the
// HOST so that its per-site arithmetic (macroscopic, equilibrium_all, kbc_collide, guo_force) can be checked
// against the oracle on a box without a GPU (tests/test_lattice_arith_cpu.py).  Built with
// g++ -O2 -ffp-contract=off: the only fused operations are the explicit fma() calls of div_const, exactly as
// in the -fmad=false CUDA build.  Nothing in the product links or loads this file.
#include <cmath>
#include <cstdint>
#include <cstring>
#define __device__
#define __host__
#define __forceinline__ inline
#include "../../luma_b200/csrc/lattice.cuh"

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

// one site: pulled populations fp (feed rho,u), own-site previous populations fo (KBC collides these)
template <class L, bool FORCE>
static void site(const double *fp, const double *fo, int kbc, double omega, const double *F3, const LbmConst &C, double *out, double *rho_out, double *u_out)
{
	double f[L::Q], own[L::Q], feq[L::Q], res[L::Q], u[3], rho;
	double F[3] = { F3[0], F3[1], F3[2] }, hF[3] = { 0.5 * F3[0], 0.5 * F3[1], 0.5 * F3[2] }, lam[4];
	for (int v = 0; v < L::Q; ++v) { f[v] = fp[v]; own[v] = fo[v]; }
	for (int k = 0; k < 4; ++k) lam[k] = (1 - 0.5 * omega) * (C.w[k] / C.cs2);
	macroscopic<L, FORCE>(f, hF, rho, u);
	equilibrium_all<L>(rho, u, C, feq);
	if (kbc)
	{
		const double beta_m1 = 2.0 / omega;
		kbc_collide<L, FORCE>(u, feq, own, beta_m1, 1.0 / beta_m1, F, C, lam, res);
	}
	else
	{
		for (int v = 0; v < L::Q; ++v)
			res[v] = FORCE ? f[v] + (omega * (feq[v] - f[v]) + guo_force<L>(v, u, F, C, lam)) : f[v] + omega * (feq[v] - f[v]);
	}
	for (int v = 0; v < L::Q; ++v) out[v] = res[v];
	*rho_out = rho;
	for (int d = 0; d < L::D; ++d) u_out[d] = u[d];
}

template <class L>
static void run(long long n, const double *fp, const double *fo, int kbc, int force, double omega, const double *F3, double *out, double *rho, double *u)
{
	LbmConst C;
	make_constants(C, L::Q);
	for (long long s = 0; s < n; ++s)
	{
		if (force) site<L, true>(fp + s * L::Q, fo + s * L::Q, kbc, omega, F3, C, out + s * L::Q, rho + s, u + s * L::D);
		else site<L, false>(fp + s * L::Q, fo + s * L::Q, kbc, omega, F3, C, out + s * L::Q, rho + s, u + s * L::D);
	}
}

extern "C" int lattice_host_sites(int Q, long long n, const double *fp, const double *fo, int kbc, int force, double omega,
	const double *F3, double *out, double *rho, double *u)
{
	if (Q == 9) run<D2Q9>(n, fp, fo, kbc, force, omega, F3, out, rho, u);
	else if (Q == 19) run<D3Q19>(n, fp, fo, kbc, force, omega, F3, out, rho, u);
	else if (Q == 27) run<D3Q27>(n, fp, fo, kbc, force, omega, F3, out, rho, u);
	else return 1;
	return 0;
}

extern "C" int lattice_host_c(int Q, int v, int d)
{
	return Q == 9 ? D2Q9::c(v, d) : (Q == 19 ? D3Q19::c(v, d) : D3Q27::c(v, d));
}
