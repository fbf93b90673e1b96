/* TEST INFRASTRUCTURE ONLY -- see luma_oracle.h.  Plain-C (gcc, no FMA contraction, no fast-math)
 * restatement of LUMA v1.7.12's level-0 time step, GridObj::LBM_multi_opt, for a serial build
 * without refinement, IBM, BFL or MPI.  Paths below are under /root/reference/LUMA/.
 *
 * The restatement keeps the reference's memory layout (AoS, inc/IVector.h:94-134), loop order
 * (i, j, k; v ascending) and the left-to-right order of every floating-point expression, because
 * parity is judged bit-for-bit.  It is pinned against the compiled reference by
 * tests/test_oracle_pinned.py.
 */
#include "luma_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* inc/Enumerations.h:84-96 */
enum { eSolid = 0, eFluid = 1, eRefined = 2, eVelocity = 6, ePressure = 7, eSlip = 8, eExtrapolateRight = 9 };
/* inc/stdafx.h:112-114 */
#define ORC_SQRT2 1.4142135623730950488016887242097
#define ORC_PI 3.14159265358979323846

/* Lattice tables, src/stdafx.cpp:81-102 (D3Q19), :114-125 (D2Q9); rest population last. */
static const int C19[19][3] = {
	{1,0,0},{-1,0,0},{0,1,0},{0,-1,0},{0,0,1},{0,0,-1},{1,1,0},{-1,-1,0},{1,-1,0},{-1,1,0},
	{0,1,1},{0,-1,-1},{0,1,-1},{0,-1,1},{1,0,1},{-1,0,-1},{-1,0,1},{1,0,-1},{0,0,0} };
/* D3Q27 (L_USE_KBC_COLLISION in 3D), src/stdafx.cpp:41-70 */
static const int C27[27][3] = {
	{1,0,0},{-1,0,0},{0,1,0},{0,-1,0},{0,0,1},{0,0,-1},{0,1,1},{0,-1,-1},{0,1,-1},{0,-1,1},
	{1,0,1},{-1,0,-1},{1,0,-1},{-1,0,1},{1,1,0},{-1,-1,0},{1,-1,0},{-1,1,0},
	{1,1,1},{-1,-1,-1},{-1,-1,1},{1,1,-1},{-1,1,1},{1,-1,-1},{1,-1,1},{-1,1,-1},{0,0,0} };
static const int C9[9][3] = {
	{1,0,0},{-1,0,0},{0,1,0},{0,-1,0},{1,1,0},{-1,-1,0},{1,-1,0},{-1,1,0},{0,0,0} };

struct OracleGrid {
	OracleCase cs_;            /* the case */
	int D, Q, N, M, K;
	int c[27][3];
	int opp[27];               /* GridUtils::dir_opposites, src/GridUtils.cpp:40-41,:54-55,:66-67 */
	double w[27];              /* src/stdafx.cpp:130-148 */
	double cs;                 /* src/stdafx.cpp:153 */
	double dh, dt, nu, omega, gravity, uref, rho_out;
	double Lx, Ly, Lz;         /* GridManager::global_edges[e?Max][0], src/GridManager.cpp:45-51 */
	int t;
	double *xpos, *ypos, *zpos;
	double *uin[3];
	int reflect[3][27];        /* GridUtils::dir_reflect, src/GridUtils.cpp:46-51,:60-64 */
	double *f, *fnew, *rho, *u, *force_xyz, *force_i;
	double *rho_timeav, *ui_timeav, *uiuj_timeav;   /* inc/GridObj.h:93-95, src/GridObj_init_grids.cpp:304-306 */
	int32_t *lattyp, *wall;
	double momex[3];
	int err;
};

#define SQ(x) ((x) * (x))    /* inc/stdafx.h:110 */

static size_t site_id(const OracleGrid *g, int i, int j, int k)
{
	return (size_t)k + (size_t)j * g->K + (size_t)i * g->K * g->M;
}

/* GridUtils::linspace, inc/GridUtils.h:289-309 */
static double *linspace(double lo, double hi, int n)
{
	if (n < 2) n = 2;
	double *r = (double *)malloc(sizeof(double) * (size_t)n);
	double spacing = (hi - lo) / (double)(n - 1);
	for (int i = 0; i < n; ++i) r[i] = lo + spacing * i;
	return r;
}

/* GridUtils::getVelocityRampCoefficient, src/GridUtils.cpp:1808-1816 */
double luma_oracle_velocity_ramp(const OracleCase *c, double t)
{
	if (c->velocity_ramp_on && t <= c->velocity_ramp)
		return (1.0 - cos(ORC_PI * t / c->velocity_ramp)) / 2.0;
	return 1.0;
}

/* GridUtils::getReynoldsRampCoefficient, src/GridUtils.cpp:1825-1833 */
double luma_oracle_reynolds_ramp(const OracleCase *c, double t)
{
	if (c->reynolds_ramp_on && t <= c->reynolds_ramp)
		return 1.0 - cos(ORC_PI * t / c->reynolds_ramp);
	return 1.0;
}

/* GridObj::LBM_setBCPrecedence, src/GridObj_init_grids.cpp:1372-1377 */
static int32_t bc_precedence(int32_t current, int32_t desired)
{
	if (current == eSolid || desired == eSolid) return eSolid;
	else if (current == eVelocity) return eVelocity;
	else return desired;
}

/* GridUtils::isWithinDomainWall, src/GridUtils.cpp:1369-1430.
 * out = {edgeCount, normalDirection, nx, ny, nz}; normalDirection 3 == eNoDirection. */
static int within_domain_wall(const OracleGrid *g, double x, double y, double z, int32_t out[5])
{
	const double *th = g->cs_.wall_thick;
	int32_t n[3] = { 0, 0, 0 }, nd = 3, ec = 0;
	if (x > 0.0 && x < th[0]) { nd = 0; n[0] = 1; ec++; }
	if (x < g->Lx && x > g->Lx - th[1]) { nd = 0; n[0] = -1; ec++; }
	if (y > 0.0 && y < th[2]) { nd = 1; n[1] = 1; ec++; }
	if (y < g->Ly && y > g->Ly - th[3]) { nd = 1; n[1] = -1; ec++; }
	if (g->D == 3)
	{
		if (z > 0.0 && z < th[4]) { nd = 2; n[2] = 1; ec++; }
		if (z < g->Lz && z > g->Lz - th[5]) { nd = 2; n[2] = -1; ec++; }
	}
	out[0] = ec; out[1] = nd; out[2] = n[0]; out[3] = n[1]; out[4] = n[2];
	return ec > 0;
}

/* GridObj::_LBM_equilibrium_opt, src/GridObj_ops_lbm_optimised.cpp:674-705 */
static double feq_at(const OracleGrid *g, size_t id, int v)
{
	const int *c = g->c[v];
	const double cs = g->cs;
	double A, B;
	if (g->D == 3)
	{
		const double *u = &g->u[id * 3];
		A = (c[0] * u[0]) + (c[1] * u[1]) + (c[2] * u[2]);
		B = (SQ(c[0]) - SQ(cs)) * SQ(u[0]) +
			(SQ(c[1]) - SQ(cs)) * SQ(u[1]) +
			(SQ(c[2]) - SQ(cs)) * SQ(u[2]) +
			2 * c[0] * c[1] * u[0] * u[1] +
			2 * c[0] * c[2] * u[0] * u[2] +
			2 * c[1] * c[2] * u[1] * u[2];
	}
	else
	{
		const double *u = &g->u[id * 2];
		A = (c[0] * u[0]) + (c[1] * u[1]);
		B = (SQ(c[0]) - SQ(cs)) * SQ(u[0]) +
			(SQ(c[1]) - SQ(cs)) * SQ(u[1]) +
			2 * c[0] * c[1] * u[0] * u[1];
	}
	return g->rho[id] * g->w[v] * (1.0 + (A / SQ(cs)) + (B / (2.0 * SQ(cs) * SQ(cs))));
}

/* GridObj::_LBM_applySpecReflect_opt, src/GridObj_ops_lbm_optimised.cpp:527-581: the first wall,
 * in the order left, right, bottom, top, front, back, whose inward normal component equals the
 * link's component reflects the link in that direction. */
static int spec_reflect(OracleGrid *g, int i, int j, int k, size_t id, int v)
{
	int32_t wd[5];
	if (!within_domain_wall(g, g->xpos[i], g->ypos[j], g->zpos[k], wd)) { g->err = 4; return 0; }
	const int32_t *n = &wd[2];
	for (int d = 0; d < 3; ++d)
	{
		if (n[d] == 1 && g->c[v][d] == 1) { g->fnew[v + id * g->Q] = g->f[g->reflect[d][v] + id * g->Q]; return 1; }
		if (n[d] == -1 && g->c[v][d] == -1) { g->fnew[v + id * g->Q] = g->f[g->reflect[d][v] + id * g->Q]; return 1; }
	}
	return 0;
}

/* GridObj::_LBM_stream_opt, src/GridObj_ops_lbm_optimised.cpp:206-297 (no BFL, refinement) */
static void stream_site(OracleGrid *g, int i, int j, int k, size_t id)
{
	const int Q = g->Q;
	const int32_t type = g->lattyp[id];
	for (int v = 0; v < Q; ++v)
	{
		int sx = (i - g->c[v][0] + g->N) % g->N;
		int sy = (j - g->c[v][1] + g->M) % g->M;
		int sz = (k - g->c[v][2] + g->K) % g->K;
		size_t src = site_id(g, sx, sy, sz);
		int32_t st = g->lattyp[src];
		if (type == eSlip)
		{
			/* slip :229-233 */
			if (spec_reflect(g, i, j, k, id, v)) continue;
			if (g->err) return;
		}
		if (st == eSolid)
		{
			/* halfway bounce-back :238-243 */
			g->fnew[v + id * Q] = g->f[g->opp[v] + id * Q];
		}
		else if (st == eExtrapolateRight)
		{
			/* the value two sites to the left of the source :246-251 */
			if (sx < 2) { g->err = 3; return; }
			g->fnew[v + id * Q] = g->f[v + (src - 2 * ((size_t)g->K * g->M)) * Q];
		}
		else if (!g->cs_.regularised && st == eVelocity)
		{
			/* forced-equilibrium velocity BC :254-270 */
			if (g->cs_.velocity_ramp_on)
			{
				double ramp = luma_oracle_velocity_ramp(&g->cs_, g->t * g->dt);
				g->u[0 + src * g->D] = g->uin[0][j] * ramp;
				g->u[1 + src * g->D] = g->uin[1][j] * ramp;
				if (g->D == 3) g->u[2 + src * g->D] = g->uin[2][j] * ramp;
			}
			g->fnew[v + id * Q] = feq_at(g, src, v);
		}
		else
		{
			g->fnew[v + id * Q] = g->f[v + src * Q];   /* :289-293 */
		}
	}
}

/* time-averaged statistics, src/GridObj_ops_lbm_optimised.cpp:895-917 */
static void time_average_site(OracleGrid *g, size_t id)
{
	const int D = g->D, P = 3 * D - 3, t = g->t;
	double ta_temp = g->rho_timeav[id] * (double)t;
	ta_temp += g->rho[id];
	g->rho_timeav[id] = ta_temp / (double)(t + 1);
	int pq_combo = 0;
	for (int p = 0; p < D; p++)
	{
		ta_temp = g->ui_timeav[p + id * D] * (double)t;
		ta_temp += g->u[p + id * D];
		g->ui_timeav[p + id * D] = ta_temp / (double)(t + 1);
		for (int q = p; q < D; q++)
		{
			ta_temp = g->uiuj_timeav[pq_combo + id * P] * (double)t;
			ta_temp += (g->u[p + id * D] * g->u[q + id * D]);
			g->uiuj_timeav[pq_combo + id * P] = ta_temp / (double)(t + 1);
			pq_combo++;
		}
	}
}

/* the moment sums of GridObj::_LBM_macro_opt, src/GridObj_ops_lbm_optimised.cpp:800-847 */
static void macro_moments(OracleGrid *g, size_t id, int32_t type)
{
	if (type != eFluid && type != eSlip) return;   /* eBFL/eTransitionToFiner never occur here */
	const int Q = g->Q, D = g->D;
	double r = 0.0, mx = 0.0, my = 0.0, mz = 0.0;
	for (int v = 0; v < Q; ++v)
	{
		double fv = g->fnew[v + id * Q];
		r += fv;
		mx += g->c[v][0] * fv;
		my += g->c[v][1] * fv;
		if (D == 3) mz += g->c[v][2] * fv;
	}
	if (g->cs_.gravity_on)
	{
		mx += 0.5 * g->force_xyz[0 + id * D];
		my += 0.5 * g->force_xyz[1 + id * D];
		if (D == 3) mx += 0.5 * g->force_xyz[2 + id * D];   /* sic, :833 adds F_z to x-momentum */
	}
	g->u[0 + id * D] = mx / r;
	g->u[1 + id * D] = my / r;
	if (D == 3) g->u[2 + id * D] = mz / r;
	g->rho[id] = r;
}

/* GridObj::_LBM_macro_opt, src/GridObj_ops_lbm_optimised.cpp:800-918 */
static void macro_site(OracleGrid *g, size_t id, int32_t type)
{
	macro_moments(g, id, type);
	if (g->cs_.time_averaged) time_average_site(g, id);
}

/* GridUtils::isOffGrid, src/GridUtils.cpp:1533-1546 */
static int off_grid(const OracleGrid *g, int i, int j, int k)
{
	return (i >= g->N || i < 0) || (j >= g->M || j < 0) || (k >= g->K || k < 0);
}

/* GridObj::_LBM_updateAndExtrapolate (order 1) :1353-1412 with _LBM_updateInteriorLatticeSite
 * :1422-1434 and GridUtils::extrapolate inc/GridUtils.h:232-280.  q is rho (stride 1, p 0) or
 * u (stride D, component p). */
static double update_and_extrapolate(OracleGrid *g, double *q, const int32_t n[3], int i, int j, int k, int p, int stride)
{
	size_t id = site_id(g, i, j, k);
	int i1 = i + n[0], i2 = i1 + n[0];
	int j1 = j + n[1], j2 = j1 + n[1];
	int k1 = k + n[2], k2 = k1 + n[2];
	if (off_grid(g, i1, j1, k1) || off_grid(g, i2, j2, k2)) { g->err = 3; return 0.0; }
	size_t id1 = site_id(g, i1, j1, k1), id2 = site_id(g, i2, j2, k2);
	if (id1 > id) { stream_site(g, i1, j1, k1, id1); macro_site(g, id1, g->lattyp[id1]); }
	if (id2 > id) { stream_site(g, i2, j2, k2, id2); macro_site(g, id2, g->lattyp[id2]); }
	return 2.0 * q[p + id1 * stride] - q[p + id2 * stride];
}

/* GridObj::_LBM_regularised_opt, src/GridObj_ops_lbm_optimised.cpp:313-510 */
static void regularised_site(OracleGrid *g, int i, int j, int k, size_t id, int32_t type)
{
	const int Q = g->Q, D = g->D;
	const double cs = g->cs;
	double uw[3] = { 0, 0, 0 };
	double dens = g->cs_.rhoin;
	int32_t wd[5];
	double f_plus = 0.0, f_zero = 0.0;
	double Sxx = 0, Syy = 0, Sxy = 0, Szz = 0, Sxz = 0, Syz = 0;
	double ramp = luma_oracle_velocity_ramp(&g->cs_, (g->t + 1) * g->dt);

	if (!within_domain_wall(g, g->xpos[i], g->ypos[j], g->zpos[k], wd)) { g->err = 1; return; }
	const int ec = wd[0], nd = wd[1];
	const int32_t *n = &wd[2];

	uw[0] = g->uin[0][j] * ramp;
	uw[1] = g->uin[1][j] * ramp;
	uw[2] = g->uin[2][j] * ramp;
	dens = g->rho_out;   /* L_RHOIN + pd2dlbm(L_PRESSURE_DELTA), :343-345 */

	if (ec > 1)
	{
		if (type == ePressure) { g->err = 2; return; }
		dens = update_and_extrapolate(g, g->rho, n, i, j, k, 0, 1);   /* :364 */
	}
	else
	{
		for (int v = 0; v < Q; ++v)
		{
			if (g->c[v][nd] == -n[nd]) f_plus += g->fnew[v + id * Q];
			else if (g->c[v][nd] == 0) f_zero += g->fnew[v + id * Q];
		}
		if (type == ePressure)
		{
			for (int d = 0; d < D; ++d)
				if (d != nd) uw[d] = update_and_extrapolate(g, g->u, n, i, j, k, d, D);   /* :397-402 */
			uw[nd] = 1.0 - ((1.0 / dens) * (2.0 * f_plus + f_zero));
			if (n[nd] == -1) uw[nd] *= -1.0;
		}
		else
		{
			double un = uw[nd];
			if (n[nd] == -1) un *= -1.0;
			dens = (1.0 / (1.0 - un)) * (2.0 * f_plus + f_zero);
		}
	}
	if (g->err) return;

	g->rho[id] = dens;
	g->u[0 + id * D] = uw[0];
	g->u[1 + id * D] = uw[1];
	if (D == 3) g->u[2 + id * D] = uw[2];

	for (int v = 0; v < Q; ++v)
	{
		const int *c = g->c[v];
		if (ec == 1 && c[nd] == n[nd])
		{
			g->fnew[v + id * Q] = feq_at(g, id, v) + (g->fnew[g->opp[v] + id * Q] - feq_at(g, id, g->opp[v]));
		}
		else if (ec > 1 && (c[0] == n[0] || c[1] == n[1] || (D == 3 && c[2] == n[2])))
		{
			int dp = 0; double mag = 0.0;
			for (int d = 0; d < D; ++d) { dp += c[d] * n[d]; mag += ((double)c[d] * (double)c[d]); }
			mag = sqrt(mag);
			if (dp == 0 && mag > 1.0)
				g->fnew[v + id * Q] = feq_at(g, id, v);   /* buried link :468-471 */
			else
				g->fnew[v + id * Q] = feq_at(g, id, v) + (g->fnew[g->opp[v] + id * Q] - feq_at(g, id, g->opp[v]));
		}
		double fneq = g->fnew[v + id * Q] - feq_at(g, id, v);
		Sxx += c[0] * c[0] * fneq;
		Syy += c[1] * c[1] * fneq;
		Sxy += c[0] * c[1] * fneq;
		if (D == 3)
		{
			Szz += c[2] * c[2] * fneq;
			Sxz += c[0] * c[2] * fneq;
			Syz += c[1] * c[2] * fneq;
		}
	}
	for (int v = 0; v < Q; ++v)
	{
		const int *c = g->c[v];
		g->fnew[v + id * Q] = feq_at(g, id, v) +
			(g->w[v] / (2.0 * SQ(cs) * SQ(cs))) *
			(
			((c[0] * c[0] - SQ(cs)) * Sxx) +
			((c[1] * c[1] - SQ(cs)) * Syy) +
			((c[2] * c[2] - SQ(cs)) * Szz) +
			(2.0 * c[0] * c[1] * Sxy) +
			(2.0 * c[0] * c[2] * Sxz) +
			(2.0 * c[1] * c[2] * Syz)
			);
	}
}

/* GridObj::_LBM_forceGrid_opt, src/GridObj_ops_lbm_optimised.cpp:928-990 */
static void force_site(OracleGrid *g, size_t id)
{
	const int Q = g->Q, D = g->D;
	const double cs = g->cs;
	memset(&g->force_i[id * Q], 0, sizeof(double) * (size_t)Q);
	for (int v = 0; v < Q; ++v)
	{
		double beta = 0.0;
		double lambda = (1 - 0.5 * g->omega) * (g->w[v] / (cs * cs));
		for (int d = 0; d < D; ++d) beta += (g->c[v][d] * g->u[d + id * D]);
		beta = beta * (1 / (cs * cs));
		for (int d = 0; d < D; ++d)
			g->force_i[v + id * Q] += g->force_xyz[d + id * D] * (g->c[v][d] * (1 + beta) - g->u[d + id * D]);
		g->force_i[v + id * Q] *= lambda;
	}
}

/* GridObj::_LBM_smag, src/GridObj_ops_lbm_optimised.cpp:717-756; Matrix2D::operator% inc/Matrix.h:65-75 */
static double smag_omega(const OracleGrid *g, size_t id, double omega)
{
	const int Q = g->Q, D = g->D;
	const double cs = g->cs;
	double S[3][3] = { { 0, 0, 0 }, { 0, 0, 0 }, { 0, 0, 0 } };
	double fneq[27];
	for (int v = 0; v < Q; ++v) fneq[v] = g->fnew[v + id * Q] - feq_at(g, id, v);
	for (int a = 0; a < D; ++a)
		for (int b = a; b < D; ++b)
		{
			S[a][b] = 0.0;
			for (int v = 0; v < Q; ++v) S[a][b] += g->c[v][a] * g->c[v][b] * fneq[v];
		}
	for (int a = 1; a < D; ++a) for (int b = 0; b < a; ++b) S[a][b] = S[b][a];
	double total = 0.0;
	for (int a = 0; a < 3; ++a)
	{
		double row = 0.0;
		for (int b = 0; b < 3; ++b) row += S[a][b] * S[a][b];
		total += row;
	}
	double Qm = sqrt(2.0 * total);
	double tau = 1.0 / omega;
	double tau_t = 0.5 * (sqrt(SQ(tau) + 2.0 * ORC_SQRT2 * SQ(g->cs_.csmag) * g->cs_.rhoin * SQ(cs) * SQ(cs) * Qm) - tau);
	return (1.0 / (tau + tau_t));
}

/* GridObj::_LBM_collide_opt, src/GridObj_ops_lbm_optimised.cpp:765-790 */
static void collide_site(OracleGrid *g, size_t id)
{
	const int Q = g->Q;
	double omega_s = g->cs_.bgksmag ? smag_omega(g, id, g->omega) : g->omega;
	for (int v = 0; v < Q; ++v)
	{
		if (g->cs_.gravity_on)
			g->fnew[v + id * Q] += omega_s * (feq_at(g, id, v) - g->fnew[v + id * Q]) + g->force_i[v + id * Q];
		else
			g->fnew[v + id * Q] += omega_s * (feq_at(g, id, v) - g->fnew[v + id * Q]);
	}
}

/* GridObj::_LBM_kbcCollide_opt, src/GridObj_ops_lbm_optimised.cpp:1122-1305 (KBC-D on D2Q9, KBC-N4 on
 * D3Q27).  NB the reference reads `f` -- the lattice of the PREVIOUS time level at this site, not the
 * streamed `fNew` (:1150, :1292) -- so the streamed (and regularised) populations only reach rho and u. */
static void kbc_collide_site(OracleGrid *g, size_t id)
{
	const int Q = g->Q, D = g->D;
	const int nm = (D == 3) ? 13 : 3;
	double ds[27], dh[27], fneq[27], feq[27], Mneq[13];
	int C[13 * 27];
	double gamma;
	for (int m = 0; m < nm; ++m) Mneq[m] = 0.0;
	for (int m = 0; m < nm * Q; ++m) C[m] = 1;
	for (int v = 0; v < Q; ++v)
	{
		feq[v] = feq_at(g, id, v);
		fneq[v] = g->f[v + id * Q] - feq[v];
		int idx = 0;
		for (int sig = 0; sig < D; ++sig)
			for (int gam = sig; gam < D; ++gam)
			{
				C[idx + v * nm] = g->c[v][sig] * g->c[v][gam];
				Mneq[idx] += fneq[v] * C[idx + v * nm];
				idx++;
				if (D == 3)
					for (int del = gam; del < D; ++del)
						if (sig != gam || gam != del || sig != del)
						{
							C[idx + v * nm] = g->c[v][sig] * g->c[v][gam] * g->c[v][del];
							Mneq[idx] += fneq[v] * C[idx + v * nm];
							idx++;
						}
			}
	}
	for (int v = 0; v < Q; ++v)
	{
		const int *c = g->c[v];
		if (D == 3)
		{
			if (c[0] == 0)
			{
				if (c[1] == 0)
				{
					if (c[2] == 0) ds[v] = (-(Mneq[0] + Mneq[8] + Mneq[12]));
					else ds[v] = ((-(Mneq[0] - Mneq[12]) - (Mneq[8] - Mneq[12])) / 6.0 + (Mneq[0] + Mneq[8] + Mneq[12]) / 6.0 - c[2] * 0.5 * (Mneq[2] + Mneq[9]));
				}
				else
				{
					if (c[2] == 0) ds[v] = ((-(Mneq[0] - Mneq[12]) + 2.0 * (Mneq[8] - Mneq[12])) / 6.0 + (Mneq[0] + Mneq[8] + Mneq[12]) / 6.0 - c[1] * 0.5 * (Mneq[1] + Mneq[11]));
					else ds[v] = (C[10 + v * nm] * 0.25 * Mneq[10] + (c[2] * 0.25 * Mneq[9] + c[1] * 0.25 * Mneq[11]));
				}
			}
			else
			{
				if (c[1] == 0)
				{
					if (c[2] == 0) ds[v] = ((2.0 * (Mneq[0] - Mneq[12]) - (Mneq[8] - Mneq[12])) / 6.0 + (Mneq[0] + Mneq[8] + Mneq[12]) / 6.0 - c[0] * 0.5 * (Mneq[4] + Mneq[7]));
					else ds[v] = (C[6 + v * nm] * 0.25 * Mneq[6] + (c[2] * 0.25 * Mneq[2] + c[0] * 0.25 * Mneq[7]));
				}
				else
				{
					if (c[2] == 0) ds[v] = (C[3 + v * nm] * 0.25 * Mneq[3] + (c[1] * 0.25 * Mneq[1] + c[0] * 0.25 * Mneq[4]));
					else ds[v] = (C[5 + v * nm] * Mneq[5] / 8.0);
				}
			}
		}
		else
		{
			if (c[0] == 0)
			{
				if (c[1] == 0) ds[v] = 0.0;
				else ds[v] = -0.25 * (Mneq[0] - Mneq[2]);
			}
			else
			{
				if (c[1] == 0) ds[v] = 0.25 * (Mneq[0] - Mneq[2]);
				else ds[v] = 0.25 * C[1 + v * nm] * Mneq[1];
			}
		}
		dh[v] = fneq[v] - ds[v];
	}
	double top_prod = 0.0, bot_prod = 0.0;
	for (int v = 0; v < Q; ++v)
	{
		top_prod += ds[v] * dh[v] / feq[v];
		bot_prod += dh[v] * dh[v] / feq[v];
	}
	double beta_m1 = 2.0 / g->omega;
	if (bot_prod == 0.0) gamma = 2.0;
	else gamma = beta_m1 - (2.0 - beta_m1) * (top_prod / bot_prod);
	for (int v = 0; v < Q; ++v)
	{
		if (g->cs_.gravity_on)
			g->fnew[v + id * Q] = g->f[v + id * Q] - (1.0 / beta_m1) * (2.0 * ds[v] + gamma * dh[v]) + g->force_i[v + id * Q];
		else
			g->fnew[v + id * Q] = g->f[v + id * Q] - (1.0 / beta_m1) * (2.0 * ds[v] + gamma * dh[v]);
	}
}

/* ObjectManager::computeLiftDrag(i,j,k,g), src/ObjectManager.cpp:93-164 */
static void momex_site(OracleGrid *g, int i, int j, int k)
{
	const int Q = g->Q;
	for (int n = 0; n < Q; ++n)
	{
		double cx = 0.0, cy = 0.0, cz = 0.0;
		int no = g->opp[n];
		int xd = i - g->c[no][0], yd = j - g->c[no][1], zd = k - g->c[no][2];
		if (!off_grid(g, xd, yd, zd) && g->lattyp[site_id(g, xd, yd, zd)] == eFluid)
		{
			double fv = g->f[no + site_id(g, xd, yd, zd) * Q];
			cx = 2.0 * g->c[no][0] * fv;
			cy = 2.0 * g->c[no][1] * fv;
			cz = 2.0 * g->c[no][2] * fv;
		}
		g->momex[0] += cx; g->momex[1] += cy; g->momex[2] += cz;
	}
}

/* GridObj::LBM_multi_opt, src/GridObj_ops_lbm_optimised.cpp:36-193 */
static void multi_opt(OracleGrid *g)
{
	const OracleCase *c = &g->cs_;
	if (c->reynolds_ramp_on)
	{
		/* _LBM_updateReynolds :1313-1321; GridUnits::nud2nulbm inc/GridUnits.h:128 */
		double newRe = (double)c->re * luma_oracle_reynolds_ramp(c, (g->t + 1) * g->dt);
		g->nu = ((1.0 / (double)newRe) * g->dt) / (SQ(g->dh));
		g->omega = 1.0 / ((g->nu / SQ(g->cs)) + 0.5);
	}
	if (c->ld_out) { g->momex[0] = g->momex[1] = g->momex[2] = 0.0; }   /* resetMomexBodyForces */

	for (int i = 0; i < g->N; ++i)
		for (int j = 0; j < g->M; ++j)
			for (int k = 0; k < g->K; ++k)
			{
				size_t id = site_id(g, i, j, k);
				int32_t type = g->lattyp[id];
				if (c->ld_out && type == eSolid) momex_site(g, i, j, k);
				if (type == eRefined || type == eSolid || (!c->regularised && type == eVelocity)) continue;
				stream_site(g, i, j, k, id);
				if (g->err) return;
				if (c->regularised && (type == eVelocity || type == ePressure))
					regularised_site(g, i, j, k, id, type);
				macro_site(g, id, type);
				if (c->gravity_on) force_site(g, id);
				if (c->kbc) kbc_collide_site(g, id);   /* :147-151 */
				else collide_site(g, id);
			}
	double *tmp = g->f; g->f = g->fnew; g->fnew = tmp;   /* f.swap(fNew) :159 */
	++g->t;
}

int luma_oracle_step(OracleGrid *g, int nsteps)
{
	for (int s = 0; s < nsteps && !g->err; ++s) multi_opt(g);
	return g->err;
}

/* GridObj::LBM_initGrid, src/GridObj_init_grids.cpp:155-384 and the helpers it calls */
OracleGrid *luma_oracle_create(const OracleCase *c)
{
	if (!c || (c->dims != 2 && c->dims != 3) || c->N < 1 || c->M < 1 || c->K < 1) return NULL;
	OracleGrid *g = (OracleGrid *)calloc(1, sizeof(OracleGrid));
	if (!g) return NULL;
	g->cs_ = *c;
	g->D = c->dims; g->Q = (c->dims == 3) ? (c->kbc ? 27 : 19) : 9;   /* definitions.h:299-310 */
	if (c->kbc && c->dims == 3 && c->regularised) { free(g); return NULL; }   /* L_ERROR, src/GridObj_init_grids.cpp:266-270 */
	g->N = c->N; g->M = c->M; g->K = (c->dims == 3) ? c->K : 1;
	const int Q = g->Q, D = g->D;
	for (int v = 0; v < Q; ++v)
		for (int d = 0; d < 3; ++d) g->c[v][d] = (D == 3) ? (Q == 27 ? C27[v][d] : C19[v][d]) : C9[v][d];
	for (int v = 0; v < Q - 1; ++v) g->opp[v] = v ^ 1;
	g->opp[Q - 1] = Q - 1;
	/* dir_reflect[plane][v]: the direction with component `plane` negated */
	for (int d = 0; d < 3; ++d)
		for (int v = 0; v < Q; ++v)
			for (int r = 0; r < Q; ++r)
			{
				int same = 1;
				for (int e = 0; e < 3; ++e) same = same && (g->c[r][e] == ((e == d) ? -g->c[v][e] : g->c[v][e]));
				if (same) g->reflect[d][v] = r;
			}
	if (Q == 27)
	{
		for (int v = 0; v < 6; ++v) g->w[v] = 2.0 / 27.0;
		for (int v = 6; v < 18; ++v) g->w[v] = 1.0 / 54.0;
		for (int v = 18; v < 26; ++v) g->w[v] = 1.0 / 216.0;
		g->w[26] = 8.0 / 27.0;
	}
	else if (D == 3)
	{
		for (int v = 0; v < 6; ++v) g->w[v] = 1.0 / 18.0;
		for (int v = 6; v < 18; ++v) g->w[v] = 1.0 / 36.0;
		g->w[18] = 1.0 / 3.0;
	}
	else
	{
		for (int v = 0; v < 4; ++v) g->w[v] = 1.0 / 9.0;
		for (int v = 4; v < 8; ++v) g->w[v] = 1.0 / 36.0;
		g->w[8] = 4.0 / 9.0;
	}
	g->cs = 1.0 / sqrt(3.0);

	/* spacing, edges (src/GridManager.cpp:42-51), unit conversions (inc/GridUnits.h) */
	g->dh = 1.0 / (double)c->resolution;
	g->dt = c->dt;
	g->Lx = g->dh * g->N; g->Ly = g->dh * g->M; g->Lz = g->dh * ((D == 3) ? g->K : 1);
	g->gravity = (c->gravity_force * SQ(g->dt)) / g->dh;
	g->uref = (1 * g->dt) / g->dh;
	{
		/* pd2dlbm, inc/GridUnits.h:178-181; dm = (L_PHYSICAL_RHO / L_RHOIN) dh^3, init_grids.cpp:178 */
		double dm = (1000.0 / c->rhoin) * g->dh * g->dh * g->dh;
		g->rho_out = c->rhoin + (c->pressure_delta * g->dh * SQ(g->dt) / dm) / SQ(g->cs);
	}

	size_t ns = (size_t)g->N * g->M * g->K;
	g->xpos = linspace(g->dh / 2.0, g->Lx - g->dh / 2.0, g->N);
	g->ypos = linspace(g->dh / 2.0, g->Ly - g->dh / 2.0, g->M);
	if (D == 3) g->zpos = linspace(g->dh / 2.0, g->Lz - g->dh / 2.0, g->K);
	else { g->zpos = (double *)malloc(2 * sizeof(double)); g->zpos[0] = 0.0; g->zpos[1] = 0.0; }
	g->lattyp = (int32_t *)malloc(ns * sizeof(int32_t));
	g->wall = (int32_t *)calloc(ns * 5, sizeof(int32_t));
	g->f = (double *)malloc(ns * Q * sizeof(double));
	g->fnew = (double *)malloc(ns * Q * sizeof(double));
	g->rho = (double *)malloc(ns * sizeof(double));
	g->u = (double *)malloc(ns * D * sizeof(double));
	g->force_xyz = (double *)calloc(ns * D, sizeof(double));
	g->force_i = (double *)calloc(ns * Q, sizeof(double));
	g->rho_timeav = (double *)calloc(ns, sizeof(double));
	g->ui_timeav = (double *)calloc(ns * D, sizeof(double));
	g->uiuj_timeav = (double *)calloc(ns * (3 * D - 3), sizeof(double));
	for (int d = 0; d < 3; ++d) g->uin[d] = (double *)calloc((size_t)g->M, sizeof(double));
	if (!g->rho_timeav || !g->ui_timeav || !g->uiuj_timeav || !g->lattyp || !g->wall || !g->f || !g->fnew || !g->rho || !g->u || !g->force_xyz || !g->force_i)
	{ luma_oracle_destroy(g); return NULL; }

	/* LBM_initBoundLab :983-1097 -- order Left, Right, Front, Back, Bottom, Top */
	for (size_t s = 0; s < ns; ++s) g->lattyp[s] = eFluid;
	const double *th = c->wall_thick;
#define LABEL_PLANE(cond, wallidx, I, J, Kk) \
	if (cond) for (int a = 0; a < (I); ++a) for (int b = 0; b < (J); ++b) { size_t s = Kk; \
		g->lattyp[s] = bc_precedence(g->lattyp[s], c->wall_type[wallidx]); }
	for (int i = 0; i < g->N; ++i) LABEL_PLANE(g->xpos[i] <= th[0], 0, g->M, g->K, site_id(g, i, a, b))
	for (int i = 0; i < g->N; ++i) LABEL_PLANE(g->xpos[i] >= g->Lx - th[1], 1, g->M, g->K, site_id(g, i, a, b))
	if (D == 3)
	{
		for (int k = 0; k < g->K; ++k) LABEL_PLANE(g->zpos[k] <= th[4], 4, g->N, g->M, site_id(g, a, b, k))
		for (int k = 0; k < g->K; ++k) LABEL_PLANE(g->zpos[k] >= g->Lz - th[5], 5, g->N, g->M, site_id(g, a, b, k))
	}
	for (int j = 0; j < g->M; ++j) LABEL_PLANE(g->ypos[j] <= th[2], 2, g->N, g->K, site_id(g, a, j, b))
	for (int j = 0; j < g->M; ++j) LABEL_PLANE(g->ypos[j] >= g->Ly - th[3], 3, g->N, g->K, site_id(g, a, j, b))
#undef LABEL_PLANE

	/* _LBM_initSetInletProfile :1322-1360 */
	if (c->parabolic_inlet)
	{
		double b = g->Ly - th[3];
		double p = (b + th[2]) / 2.0;
		double q = b - p;
		for (int j = 0; j < g->M; ++j)
		{
			g->uin[0][j] = ((1.5 * c->u0[0] * g->dt) / g->dh) * (1.0 - pow((g->ypos[j] - p) / q, 2.0));
			g->uin[1][j] = 0.0;
			g->uin[2][j] = 0.0;
		}
	}
	else
	{
		for (int j = 0; j < g->M; ++j)
			for (int d = 0; d < 3; ++d) g->uin[d][j] = (c->u0[d] * g->dt) / g->dh;
		if (D == 2) for (int j = 0; j < g->M; ++j) g->uin[2][j] = (0.0 * g->dt) / g->dh;   /* L_UZ0 0.0 in 2D */
	}

	/* LBM_initVelocity :37-133, LBM_initRho :138-150 */
	double ramp0 = luma_oracle_velocity_ramp(c, 0.0);
	for (int i = 0; i < g->N; ++i) for (int j = 0; j < g->M; ++j) for (int k = 0; k < g->K; ++k)
	{
		size_t id = site_id(g, i, j, k);
		if (c->no_flow && g->lattyp[id] != eVelocity)
		{
			for (int d = 0; d < D; ++d) g->u[d + id * D] = 0.0;
		}
		else
		{
			for (int d = 0; d < D; ++d) g->u[d + id * D] = g->uin[d][j] * ramp0;
		}
		if (g->lattyp[id] == eSolid) for (int d = 0; d < D; ++d) g->u[d + id * D] = 0.0;
		g->rho[id] = c->rhoin;
	}
	if (c->gravity_on)
		for (size_t id = 0; id < ns; ++id) g->force_xyz[c->gravity_dir + id * D] = g->rho[id] * g->gravity * 1.0;

	for (size_t id = 0; id < ns; ++id)
		for (int v = 0; v < Q; ++v) g->f[v + id * Q] = feq_at(g, id, v);
	memcpy(g->fnew, g->f, ns * Q * sizeof(double));

	if (c->use_nu) g->nu = (c->nu * g->dt) / (SQ(g->dh));
	else g->nu = ((1.0 / (double)c->re) * g->dt) / (SQ(g->dh));
	g->omega = 1.0 / ((g->nu / SQ(g->cs)) + 0.5);
	if (!c->bgksmag && g->omega >= 2.0) { luma_oracle_destroy(g); return NULL; }   /* :353-356 */

	/* body: ObjectManager::addBouncebackObject(g, geom, pts), src/ObjectManager.cpp:309-345 */
	if (c->has_box)
		for (int i = c->box[0]; i < c->box[1]; ++i) for (int j = c->box[2]; j < c->box[3]; ++j)
			for (int k = c->box[4]; k < c->box[5]; ++k)
			{
				if (off_grid(g, i, j, k)) continue;
				size_t id = site_id(g, i, j, k);
				if (g->lattyp[id] == eFluid)
				{
					g->lattyp[id] = eSolid;
					g->u[0 + id * D] = 0.0;
					g->u[1 + id * D] = 0.0;   /* the reference zeroes component 0 twice in 3-D, never component 2 */
					g->rho[id] = c->rhoin;
				}
			}

	for (int i = 0; i < g->N; ++i) for (int j = 0; j < g->M; ++j) for (int k = 0; k < g->K; ++k)
		within_domain_wall(g, g->xpos[i], g->ypos[j], g->zpos[k], &g->wall[site_id(g, i, j, k) * 5]);
	for (size_t id = 0; id < ns; ++id) if (g->wall[id * 5] == 0) g->wall[id * 5 + 1] = 0;
	return g;
}

void luma_oracle_destroy(OracleGrid *g)
{
	if (!g) return;
	free(g->xpos); free(g->ypos); free(g->zpos);
	for (int d = 0; d < 3; ++d) free(g->uin[d]);
	free(g->f); free(g->fnew); free(g->rho); free(g->u); free(g->force_xyz); free(g->force_i);
	free(g->rho_timeav); free(g->ui_timeav); free(g->uiuj_timeav);
	free(g->lattyp); free(g->wall);
	free(g);
}

double  *luma_oracle_f(OracleGrid *g) { return g->f; }
double  *luma_oracle_fnew(OracleGrid *g) { return g->fnew; }
double  *luma_oracle_rho(OracleGrid *g) { return g->rho; }
double  *luma_oracle_u(OracleGrid *g) { return g->u; }
double  *luma_oracle_rho_timeav(OracleGrid *g) { return g->rho_timeav; }
double  *luma_oracle_ui_timeav(OracleGrid *g) { return g->ui_timeav; }
double  *luma_oracle_uiuj_timeav(OracleGrid *g) { return g->uiuj_timeav; }
int32_t *luma_oracle_lattyp(OracleGrid *g) { return g->lattyp; }
int32_t *luma_oracle_wall(OracleGrid *g) { return g->wall; }
double  *luma_oracle_uin(OracleGrid *g, int d) { return g->uin[d]; }
double  *luma_oracle_pos(OracleGrid *g, int d) { return d == 0 ? g->xpos : (d == 1 ? g->ypos : g->zpos); }
double   luma_oracle_omega(const OracleGrid *g) { return g->omega; }
double   luma_oracle_nu(const OracleGrid *g) { return g->nu; }
double   luma_oracle_gravity(const OracleGrid *g) { return g->gravity; }
double   luma_oracle_rho_out(const OracleGrid *g) { return g->rho_out; }
int      luma_oracle_t(const OracleGrid *g) { return g->t; }
void     luma_oracle_force(const OracleGrid *g, double F[3]) { F[0] = g->momex[0]; F[1] = g->momex[1]; F[2] = g->momex[2]; }
