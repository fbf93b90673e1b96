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

// one site: pulled populations fp (feed rho,u), own-site previous populations fo (KBC collides these).
// FORCE = 0 none, 1 + gravity direction; Fg = the one non-zero component of force_xyz
template <class L, int FORCE>
static void site(const double *fp, const double *fo, int kbc, double omega, double Fg, const LbmConst &C, double *out, double *rho_out, double *u_out)
{
	double f[L::Q], own[L::Q], feq[L::Q], res[L::Q], u[3], rho;
	double lam[4];
	for (int v = 0; v < L::Q; ++v) { f[v] = fp[v]; own[v] = fo[v]; }
	for (int k = 0; k < 4; ++k) lam[k] = (1 - 0.5 * omega) * (C.w[k] / C.cs2);
	macroscopic<L, FORCE>(f, 0.5 * Fg, rho, u);
	equilibrium_all<L>(rho, u, C, feq);
	if (kbc)
	{
		const double beta_m1 = 2.0 / omega;
		kbc_collide<L, FORCE>(u, feq, own, beta_m1, 1.0 / beta_m1, Fg, C, lam, res);
	}
	else
	{
		for (int v = 0; v < L::Q; ++v)
		{
			if constexpr (FORCE != 0) res[v] = f[v] + (omega * (feq[v] - f[v]) + guo_force<L, (FORCE > 0 ? FORCE - 1 : 0)>(v, u, Fg, C, lam));
			else res[v] = f[v] + omega * (feq[v] - f[v]);
		}
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
	int g = 0;
	for (int d = 0; d < 3; ++d) if (F3[d] != 0.0) g = d;
	const int code = force ? 1 + g : 0;
	for (long long s = 0; s < n; ++s)
	{
		const double *a = fp + s * L::Q, *b = fo + s * L::Q;
		double *o = out + s * L::Q, *r = rho + s, *uu = u + s * L::D;
		switch (code)
		{
		case 0: site<L, 0>(a, b, kbc, omega, 0.0, C, o, r, uu); break;
		case 1: site<L, 1>(a, b, kbc, omega, F3[0], C, o, r, uu); break;
		case 2: site<L, 2>(a, b, kbc, omega, F3[1], C, o, r, uu); break;
		default: if constexpr (L::D == 3) site<L, 3>(a, b, kbc, omega, F3[2], C, o, r, uu); break;
		}
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
