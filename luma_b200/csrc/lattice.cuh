// lattice.cuh -- lattice tables and the exact (reference-order) fp64 arithmetic of the LBM step.
//
// Everything here restates arithmetic of /root/reference/LUMA/src/GridObj_ops_lbm_optimised.cpp so
// that the CUDA path is bit-identical with the reference's CPU build (g++ -O3, x86-64 baseline:
// IEEE double, no FMA contraction, no reassociation).  Rules kept throughout:
//   * every sum is evaluated left to right in the reference's term order; terms the reference
//     multiplies by an integer 0 are dropped (adding +-0 never changes a running IEEE sum that is
//     later added to a non-zero value); multiplications by +-1 / +-2 are exact sign/exponent changes;
//   * this translation unit is compiled with -fmad=false: no contraction anywhere, the only fused
//     operations are the explicit fma() calls in div_const();
//   * x / cs^2 and x / (2 cs^4) are divisions by run-time CONSTANTS; div_const() evaluates them with
//     one multiply and two FMAs and is correctly rounded for these two divisors for every normal x
//     (proof by enumeration of the candidate failures: tests/test_constdiv_exact.py), i.e. it
//     returns exactly what the reference's `/` returns.
#pragma once
#include <cstdint>

namespace luma {

// ---- direction tables: src/stdafx.cpp:81-102 (D3Q19), :114-125 (D2Q9), :41-70 (D3Q27, the lattice of
//      L_USE_KBC_COLLISION builds in 3D); rest population last; opposites are (v ^ 1) for v < Q-1
//      (src/GridUtils.cpp:40-41, :54-55, :66-67).
//      WREST = weight class of the rest population; DESC_IN_WORD = the wall descriptor fits the cell word
//      beside the link bits; REGULARISABLE = L_REGULARISED_BOUNDARIES is allowed on this lattice (the
//      reference refuses it on D3Q27, src/GridObj_init_grids.cpp:266-270) ----
struct D3Q19
{
	static constexpr int Q = 19;
	static constexpr int D = 3;
	static constexpr int WREST = 2;
	static constexpr bool DESC_IN_WORD = true;
	static constexpr bool REGULARISABLE = true;
	__host__ __device__ static constexpr int c(int v, int d)
	{
		constexpr int T[19][3] = {
			{ 1, 0, 0 }, { -1, 0, 0 }, { 0, 1, 0 }, { 0, -1, 0 }, { 0, 0, 1 }, { 0, 0, -1 },
			{ 1, 1, 0 }, { -1, -1, 0 }, { 1, -1, 0 }, { -1, 1, 0 },
			{ 0, 1, 1 }, { 0, -1, -1 }, { 0, 1, -1 }, { 0, -1, 1 },
			{ 1, 0, 1 }, { -1, 0, -1 }, { -1, 0, 1 }, { 1, 0, -1 }, { 0, 0, 0 } };
		return T[v][d];
	}
	// weight class: 0 -> 1/18, 1 -> 1/36, 2 -> 1/3   (src/stdafx.cpp:140-143)
	__host__ __device__ static constexpr int wclass(int v) { return v < 6 ? 0 : (v < 18 ? 1 : 2); }
};

struct D2Q9
{
	static constexpr int Q = 9;
	static constexpr int D = 2;
	static constexpr int WREST = 2;
	static constexpr bool DESC_IN_WORD = true;
	static constexpr bool REGULARISABLE = true;
	__host__ __device__ static constexpr int c(int v, int d)
	{
		constexpr int T[9][3] = {
			{ 1, 0, 0 }, { -1, 0, 0 }, { 0, 1, 0 }, { 0, -1, 0 },
			{ 1, 1, 0 }, { -1, -1, 0 }, { 1, -1, 0 }, { -1, 1, 0 }, { 0, 0, 0 } };
		return T[v][d];
	}
	// weight class: 0 -> 1/9, 1 -> 1/36, 2 -> 4/9   (src/stdafx.cpp:147-148)
	__host__ __device__ static constexpr int wclass(int v) { return v < 4 ? 0 : (v < 8 ? 1 : 2); }
};

struct D3Q27
{
	static constexpr int Q = 27;
	static constexpr int D = 3;
	static constexpr int WREST = 3;
	static constexpr bool DESC_IN_WORD = false;
	static constexpr bool REGULARISABLE = false;
	__host__ __device__ static constexpr int c(int v, int d)
	{
		constexpr int T[27][3] = {
			{ 1, 0, 0 }, { -1, 0, 0 }, { 0, 1, 0 }, { 0, -1, 0 }, { 0, 0, 1 }, { 0, 0, -1 },
			{ 0, 1, 1 }, { 0, -1, -1 }, { 0, 1, -1 }, { 0, -1, 1 },
			{ 1, 0, 1 }, { -1, 0, -1 }, { 1, 0, -1 }, { -1, 0, 1 },
			{ 1, 1, 0 }, { -1, -1, 0 }, { 1, -1, 0 }, { -1, 1, 0 },
			{ 1, 1, 1 }, { -1, -1, -1 }, { -1, -1, 1 }, { 1, 1, -1 }, { -1, 1, 1 }, { 1, -1, -1 }, { 1, -1, 1 }, { -1, 1, -1 },
			{ 0, 0, 0 } };
		return T[v][d];
	}
	// weight class: 0 -> 2/27, 1 -> 1/54, 2 -> 1/216, 3 -> 8/27   (src/stdafx.cpp:130-136)
	__host__ __device__ static constexpr int wclass(int v) { return v < 6 ? 0 : (v < 18 ? 1 : (v < 26 ? 2 : 3)); }
};

template <class L> __host__ __device__ constexpr int opposite(int v) { return v == L::Q - 1 ? v : (v ^ 1); }

// GridUtils::getReflect (src/GridUtils.cpp:46-51, :60-64, :495): the direction whose component
// `plane` is negated and whose other components are those of v
template <class L> __host__ __device__ constexpr int reflect(int v, int plane)
{
	for (int r = 0; r < L::Q; ++r)
	{
		bool same = true;
		for (int e = 0; e < 3; ++e) same = same && (L::c(r, e) == ((e == plane) ? -L::c(v, e) : L::c(v, e)));
		if (same) return r;
	}
	return v;
}

// ---- constants derived on the host exactly as the reference derives them ----
struct LbmConst
{
	double cs2;       // SQ(cs), cs = 1.0/sqrt(3.0)                         src/stdafx.cpp:153
	double inv_cs2;   // 1.0 / cs2   (correctly rounded, host division)
	double den;       // (2.0 * cs2) * cs2                                   optimised.cpp:704
	double inv_den;   // 1.0 / den
	double k1;        // 1.0 - cs2   = (SQ(c) - SQ(cs)) for c = +-1
	double k0;        // 0.0 - cs2   for c = 0
	double w[4];      // lattice weights by class
	double wden[4];   // w[cls] / den                                        optimised.cpp:498
};

// correctly rounded a / b for the two constant divisors (y = RN(1/b)); see header comment
__device__ __forceinline__ double div_const(double a, double b, double y)
{
	const double q = a * y;
	const double r = fma(-q, b, a);
	return fma(r, y, q);
}

// ---- cell word (one uint32 per site, built once by k_cell_words); D2Q9 / D3Q19 layout ----
//  bits  0..17  link v (v < Q-1) bounces back: the site this population is pulled from is eSolid  (optimised.cpp:238)
//  bit   18     site lies on the first/last row or column of the array (periodic wrap needed in y or z)
//  bits 19..21  class: 0 not updated (eSolid, eRefined, non-regularised eVelocity), 1 eFluid with ordinary
//               sources (the k_step fast path), 2 regularised eVelocity, 3 regularised ePressure, 4 general
//               (eSlip, eExtrapolateRight, non-regularised ePressure, and eFluid sites that pull from an
//               eExtrapolateRight / forced-equilibrium eVelocity source): handled per link from the eType array
//  bits 22..23  normalDirection          } wall descriptor of velocity/pressure/slip sites,
//  bits 24..29  normal vector + 1 (2b x3)} GridUtils::isWithinDomainWall, src/GridUtils.cpp:1369
//  bits 30..31  edgeCount                }
//  D3Q27 has 26 links: bits 0..25 links, 26 the first/last row or column flag, 27..29 the class; the wall
//  descriptor (only slip sites need it there: no regularised boundaries on D3Q27) stays in the separate
//  per-site descriptor array, with the same bit positions 22..31.
enum : uint32_t { CW_CLASS_MASK = 7u, CW_ND_SHIFT = 22, CW_N_SHIFT = 24, CW_EC_SHIFT = 30 };
template <class L> struct CW
{
	static constexpr uint32_t LINKS = (L::Q == 27) ? 0x3FFFFFFu : 0x3FFFFu;
	static constexpr uint32_t EDGE = (L::Q == 27) ? (1u << 26) : (1u << 18);
	static constexpr int CLASS_SHIFT = (L::Q == 27) ? 27 : 19;
};
enum : uint32_t { CLS_SKIP = 0, CLS_FLUID = 1, CLS_VELOCITY = 2, CLS_PRESSURE = 3, CLS_GENERAL = 4 };
// eType, inc/Enumerations.h:84-96
enum : uint8_t { T_SOLID = 0, T_FLUID = 1, T_REFINED = 2, T_VELOCITY = 6, T_PRESSURE = 7, T_SLIP = 8, T_EXTRAPOLATE_RIGHT = 9 };

template <class L> __host__ __device__ inline uint32_t cw_class(uint32_t w) { return (w >> CW<L>::CLASS_SHIFT) & CW_CLASS_MASK; }
__host__ __device__ inline uint32_t cw_pack_bc(int ec, int nd, int nx, int ny, int nz)
{
	return ((uint32_t)(ec & 3) << CW_EC_SHIFT) | ((uint32_t)(nd & 3) << CW_ND_SHIFT) |
		((uint32_t)((nx + 1) | ((ny + 1) << 2) | ((nz + 1) << 4)) << CW_N_SHIFT);
}

// ---- c_v . u in the reference's term order (d = 0, 1, 2), components with c = 0 dropped.  Used for A of the
//      equilibrium and for beta of the Guo force: one expression tree, so the compiler keeps one copy.  For the
//      opposite direction the value is the exact negative (IEEE rounding is sign-symmetric). ----
template <class L>
__device__ __forceinline__ double dir_dot(const int v, const double (&u)[3])
{
	const int c0 = L::c(v, 0), c1 = L::c(v, 1), c2 = L::c(v, 2);
	double A = 0.0;
	bool first = true;
	if (c0 != 0) { A = (c0 > 0) ? u[0] : -u[0]; first = false; }
	if (c1 != 0) { const double t = (c1 > 0) ? u[1] : -u[1]; A = first ? t : A + t; first = false; }
	if (L::D == 3 && c2 != 0) { const double t = (c2 > 0) ? u[2] : -u[2]; A = first ? t : A + t; first = false; }
	return A;
}

// ---- equilibrium for all Q directions, GridObj::_LBM_equilibrium_opt (optimised.cpp:674-705):
//        feq = rho * w[v] * (1.0 + (A / SQ(cs)) + (B / (2.0 * SQ(cs) * SQ(cs))))
//      evaluated per opposite pair: A(opp) = -A and B(opp) = B hold exactly in IEEE arithmetic. ----
template <class L>
__device__ __forceinline__ void equilibrium_all(const double rho, const double (&u)[3], const LbmConst &C, double (&feq)[L::Q])
{
	double t1[3], t0[3];
#pragma unroll
	for (int d = 0; d < L::D; ++d)
	{
		const double s = u[d] * u[d];
		t1[d] = C.k1 * s;
		t0[d] = C.k0 * s;
	}
	// 2*ca*cb*ua*ub evaluates as ((+-2) * ua) * ub
	const double x01 = (2.0 * u[0]) * u[1];
	const double x02 = (L::D == 3) ? (2.0 * u[0]) * u[2] : 0.0;
	const double x12 = (L::D == 3) ? (2.0 * u[1]) * u[2] : 0.0;
	double rw[4];
#pragma unroll
	for (int k = 0; k <= L::WREST; ++k) rw[k] = rho * C.w[k];

#pragma unroll
	for (int v = 0; v < L::Q - 1; v += 2)
	{
		const int c0 = L::c(v, 0), c1 = L::c(v, 1), c2 = L::c(v, 2);
		const double A = dir_dot<L>(v, u);
		double B = (c0 ? t1[0] : t0[0]) + (c1 ? t1[1] : t0[1]);
		if (L::D == 3) B = B + (c2 ? t1[2] : t0[2]);
		if (c0 * c1 != 0) B = B + ((c0 * c1 > 0) ? x01 : -x01);
		if (L::D == 3 && c0 * c2 != 0) B = B + ((c0 * c2 > 0) ? x02 : -x02);
		if (L::D == 3 && c1 * c2 != 0) B = B + ((c1 * c2 > 0) ? x12 : -x12);
		const double qa = div_const(A, C.cs2, C.inv_cs2);
		const double qb = div_const(B, C.den, C.inv_den);
		const double r = rw[L::wclass(v)];
		feq[v] = r * ((1.0 + qa) + qb);
		feq[v + 1] = r * ((1.0 - qa) + qb);
	}
	{
		double B = t0[0] + t0[1];
		if (L::D == 3) B = B + t0[2];
		const double qb = div_const(B, C.den, C.inv_den);
		feq[L::Q - 1] = rw[L::WREST] * (1.0 + qb);
	}
}

// ---- rho = sum f, rho*u = sum c f (+ F/2), GridObj::_LBM_macro_opt (optimised.cpp:800-847).
//      FORCE = 0: no body force; FORCE = 1 + L_GRAVITY_DIRECTION otherwise.  force_xyz is uniform and has ONE non-zero
//      component (rho_init * gravity along the gravity direction, src/GridObj_init_grids.cpp:296-297), so the other
//      two additions of the reference add +0.0 to a sum that is never -0.0 and are dropped. ----
template <class L, int FORCE>
__device__ __forceinline__ void macroscopic(const double (&f)[L::Q], const double hFg, double &rho, double (&u)[3])
{
	double r = f[0];
	double m[3] = { 0.0, 0.0, 0.0 };
	bool first[3] = { true, true, true };
#pragma unroll
	for (int v = 0; v < L::Q; ++v)
	{
		if (v > 0) r = r + f[v];
#pragma unroll
		for (int d = 0; d < L::D; ++d)
		{
			const int c = L::c(v, d);
			if (c != 0)
			{
				const double t = (c > 0) ? f[v] : -f[v];
				m[d] = first[d] ? t : m[d] + t;
				first[d] = false;
			}
		}
	}
	if (FORCE == 1) m[0] = m[0] + hFg;
	if (FORCE == 2) m[1] = m[1] + hFg;
	if (FORCE == 3 && L::D == 3) m[0] = m[0] + hFg;   // sic: the reference adds F_z/2 to the x momentum (optimised.cpp:833)
	rho = r;
	u[0] = m[0] / r;
	u[1] = m[1] / r;
	u[2] = (L::D == 3) ? m[2] / r : 0.0;
}

// ---- Smagorinsky relaxation, GridObj::_LBM_smag (optimised.cpp:717-756) with Matrix2D::operator%
//      (inc/Matrix.h:65-75): returns omega_s ----
template <class L>
__device__ __forceinline__ double smagorinsky_omega(const double (&f)[L::Q], const double (&feq)[L::Q], const double tau, const double smag_coef)
{
	double S[3][3] = { { 0.0, 0.0, 0.0 }, { 0.0, 0.0, 0.0 }, { 0.0, 0.0, 0.0 } };
	bool first[3][3] = { { true, true, true }, { true, true, true }, { true, true, true } };
#pragma unroll
	for (int v = 0; v < L::Q; ++v)
	{
		const double fneq = f[v] - feq[v];
#pragma unroll
		for (int a = 0; a < L::D; ++a)
#pragma unroll
			for (int b = a; b < L::D; ++b)
			{
				const int cc = L::c(v, a) * L::c(v, b);
				if (cc != 0)
				{
					const double t = (cc > 0) ? fneq : -fneq;
					S[a][b] = first[a][b] ? t : S[a][b] + t;
					first[a][b] = false;
				}
			}
	}
#pragma unroll
	for (int a = 1; a < L::D; ++a)
#pragma unroll
		for (int b = 0; b < a; ++b) S[a][b] = S[b][a];
	double total = 0.0;
#pragma unroll
	for (int a = 0; a < L::D; ++a)
	{
		double row = S[a][0] * S[a][0];
#pragma unroll
		for (int b = 1; b < L::D; ++b) row = row + S[a][b] * S[a][b];
		total = (a == 0) ? row : total + row;
	}
	const double Qm = sqrt(2.0 * total);
	const double tau_t = 0.5 * (sqrt((tau * tau) + smag_coef * Qm) - tau);
	return 1.0 / (tau + tau_t);
}

// ---- Guo forcing term of one direction, GridObj::_LBM_forceGrid_opt (optimised.cpp:959-989):
//        beta = (sum_d c_d u_d) * (1/cs^2);   F_v = (sum_d F_d * (c_d * (1 + beta) - u_d)) * lambda_v
//      G = L_GRAVITY_DIRECTION, the one direction in which force_xyz is non-zero (Fg); the terms of the other
//      directions are (+-0) added to a running sum and cannot change it, so
//        F_v = (Fg * (c_G * (1 + beta) - u_G)) * lambda_v
//      with c_G * (1 + beta) = +-(1 + beta) or +0.  beta of the odd member of an opposite pair is the exact negative
//      of the even member's, 1.0 + (-b) == 1.0 - b; directions with c_G = 0 share one value per weight class
//      (the compiler folds the repeated expressions of the unrolled loop).  Was: 2 dot products per direction. ----
template <class L, int G>
__device__ __forceinline__ double guo_force(const int v, const double (&u)[3], const double Fg, const LbmConst &C, const double (&lam)[4])
{
	const int cg = L::c(v, G);
	double term;
	if (cg == 0) term = 0.0 - u[G];
	else
	{
		const int ve = v & ~1;      // cg != 0 rules out the rest population: ve, ve + 1 are an opposite pair
		const double b = dir_dot<L>(ve, u) * C.inv_cs2;
		const double p = (v == ve) ? 1.0 + b : 1.0 - b;
		term = (cg > 0) ? p - u[G] : (-p) - u[G];
	}
	return (Fg * term) * lam[L::wclass(v)];
}

// ---- KBC collision, GridObj::_LBM_kbcCollide_opt (optimised.cpp:1122-1305): KBC-D on D2Q9, KBC-N4 on
//      D3Q27.  `fo` are the populations the reference reads there: `f`, the PREVIOUS time level at this
//      very site (:1150, :1292) -- not the streamed fNew, which only feeds rho and u.  Moment order
//      (:1153-1176): 2-D xx, xy, yy; 3-D xx, xxy, xxz, xy, xyy, xyz, xz, xzz, yy, yyz, yz, yzz, zz. ----
template <class L> __host__ __device__ constexpr int kbc_coef(int v, int m)
{
	constexpr int T3[13][3] = { { 0, 0, -1 }, { 0, 0, 1 }, { 0, 0, 2 }, { 0, 1, -1 }, { 0, 1, 1 }, { 0, 1, 2 }, { 0, 2, -1 },
		{ 0, 2, 2 }, { 1, 1, -1 }, { 1, 1, 2 }, { 1, 2, -1 }, { 1, 2, 2 }, { 2, 2, -1 } };
	constexpr int T2[3][3] = { { 0, 0, -1 }, { 0, 1, -1 }, { 1, 1, -1 } };
	const int a = (L::D == 3) ? T3[m][0] : T2[m][0], b = (L::D == 3) ? T3[m][1] : T2[m][1], t = (L::D == 3) ? T3[m][2] : T2[m][2];
	return L::c(v, a) * L::c(v, b) * (t < 0 ? 1 : L::c(v, t));
}

template <class L, int FORCE>
__device__ __forceinline__ void kbc_collide(const double (&u)[3], const double (&feq)[L::Q], const double (&fo)[L::Q],
	const double beta_m1, const double inv_beta, const double Fg, const LbmConst &C, const double (&lam)[4], double (&out)[L::Q])
{
	constexpr int NM = (L::D == 3) ? 13 : 3;
	double M[NM], fneq[L::Q], ds[L::Q], dh[L::Q];
#pragma unroll
	for (int v = 0; v < L::Q; ++v) fneq[v] = fo[v] - feq[v];
#pragma unroll
	for (int m = 0; m < NM; ++m)
	{
		double acc = 0.0;
#pragma unroll
		for (int v = 0; v < L::Q; ++v)
		{
			const int cf = kbc_coef<L>(v, m);
			if (cf != 0) acc = acc + ((cf > 0) ? fneq[v] : -fneq[v]);
		}
		M[m] = acc;
	}
#pragma unroll
	for (int v = 0; v < L::Q; ++v)
	{
		const int c0 = L::c(v, 0), c1 = L::c(v, 1), c2 = L::c(v, 2);
		double d;
		if constexpr (L::D == 3)
		{
			// index 0 xx, 1 xxy, 2 xxz, 3 xy, 4 xyy, 5 xyz, 6 xz, 7 xzz, 8 yy, 9 yyz, 10 yz, 11 yzz, 12 zz
			const double m0 = M[0], m1 = M[1], m2 = M[2], m3 = M[3], m4 = M[4], m5 = M[5], m6 = M[6],
				m7 = M[7], m8 = M[8], m9 = M[9], m10 = M[10], m11 = M[11], m12 = M[12];
			if (c0 == 0)
			{
				if (c1 == 0)
				{
					if (c2 == 0) d = (-(m0 + m8 + m12));
					else d = ((-(m0 - m12) - (m8 - m12)) / 6.0 + (m0 + m8 + m12) / 6.0 - (double)c2 * 0.5 * (m2 + m9));
				}
				else
				{
					if (c2 == 0) d = ((-(m0 - m12) + 2.0 * (m8 - m12)) / 6.0 + (m0 + m8 + m12) / 6.0 - (double)c1 * 0.5 * (m1 + m11));
					else d = ((double)(c1 * c2) * 0.25 * m10 + ((double)c2 * 0.25 * m9 + (double)c1 * 0.25 * m11));
				}
			}
			else
			{
				if (c1 == 0)
				{
					if (c2 == 0) d = ((2.0 * (m0 - m12) - (m8 - m12)) / 6.0 + (m0 + m8 + m12) / 6.0 - (double)c0 * 0.5 * (m4 + m7));
					else d = ((double)(c0 * c2) * 0.25 * m6 + ((double)c2 * 0.25 * m2 + (double)c0 * 0.25 * m7));
				}
				else
				{
					if (c2 == 0) d = ((double)(c0 * c1) * 0.25 * m3 + ((double)c1 * 0.25 * m1 + (double)c0 * 0.25 * m4));
					else d = ((double)(c0 * c1 * c2) * m5 / 8.0);
				}
			}
		}
		else
		{
			if (c0 == 0)
			{
				if (c1 == 0) d = 0.0;
				else d = -0.25 * (M[0] - M[2]);
			}
			else
			{
				if (c1 == 0) d = 0.25 * (M[0] - M[2]);
				else d = 0.25 * (double)(c0 * c1) * M[1];
			}
		}
		ds[v] = d;
		dh[v] = fneq[v] - d;
	}
	double top = 0.0, bot = 0.0;
#pragma unroll
	for (int v = 0; v < L::Q; ++v)
	{
		top = top + ds[v] * dh[v] / feq[v];
		bot = bot + dh[v] * dh[v] / feq[v];
	}
	double gamma = 2.0;
	if (!(bot == 0.0)) gamma = beta_m1 - (2.0 - beta_m1) * (top / bot);
#pragma unroll
	for (int v = 0; v < L::Q; ++v)
	{
		if constexpr (FORCE != 0) out[v] = fo[v] - inv_beta * (2.0 * ds[v] + gamma * dh[v]) + guo_force<L, (FORCE > 0 ? FORCE - 1 : 0)>(v, u, Fg, C, lam);
		else out[v] = fo[v] - inv_beta * (2.0 * ds[v] + gamma * dh[v]);
	}
}

}  // namespace luma
