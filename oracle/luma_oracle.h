/* TEST INFRASTRUCTURE ONLY -- CPU restatement of LUMA's level-0 time step used as the parity
 * oracle.  Nothing in the product (luma_b200/, include/) may include, link or call this.
 * Allowed users: tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference legs.
 *
 * PARITY IS PINNED: the port is checked bit-for-bit against the unmodified reference sources compiled
 * as oracle/_ref/luma_ref_<case> (tests/test_oracle_pinned.py; digests of the reference's outputs are
 * committed under tests/golden/ by tests/golden/make_golden.py).  The reference's own test-suite
 * fixtures (cases/testsuite) cannot pin this path: they are v1.2.0-alpha, KBC/IBM only.
 */
#ifndef LUMA_ORACLE_H
#define LUMA_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Run-time image of the compile-time case macros of inc/definitions.h. */
typedef struct OracleCase {
	int32_t dims;               /* L_DIMS */
	int32_t resolution;         /* L_RESOLUTION */
	int32_t N, M, K;            /* L_N, L_M, L_K (definitions.h:46-48, :320-321) */
	double  dt;                 /* L_TIMESTEP */
	int32_t wall_type[6];       /* L_WALL_LEFT, RIGHT, BOTTOM, TOP, FRONT, BACK (eType values) */
	double  wall_thick[6];      /* L_WALL_THICKNESS_* in the same order (dimensionless) */
	double  u0[3];              /* L_UX0, L_UY0, L_UZ0 (dimensionless) */
	double  rhoin;              /* L_RHOIN */
	int32_t use_nu;             /* L_NU defined */
	double  nu;                 /* L_NU */
	double  re;                 /* L_RE */
	int32_t regularised;        /* L_REGULARISED_BOUNDARIES */
	int32_t no_flow;            /* L_NO_FLOW */
	int32_t bgksmag;            /* L_USE_BGKSMAG */
	double  csmag;              /* L_CSMAG */
	int32_t gravity_on;         /* L_GRAVITY_ON */
	double  gravity_force;      /* L_GRAVITY_FORCE */
	int32_t gravity_dir;        /* L_GRAVITY_DIRECTION */
	int32_t velocity_ramp_on;   /* L_VELOCITY_RAMP defined */
	double  velocity_ramp;      /* L_VELOCITY_RAMP */
	int32_t reynolds_ramp_on;   /* L_REYNOLDS_RAMP defined */
	double  reynolds_ramp;      /* L_REYNOLDS_RAMP */
	int32_t parabolic_inlet;    /* L_PARABOLIC_INLET */
	double  pressure_delta;     /* L_PRESSURE_DELTA */
	int32_t ld_out;             /* L_LD_OUT */
	int32_t has_box;            /* bounce-back body given as an index box */
	int32_t box[6];             /* i0,i1,j0,j1,k0,k1 (half-open) */
	int32_t time_averaged;      /* L_COMPUTE_TIME_AVERAGED_QUANTITIES */
	int32_t kbc;                /* L_USE_KBC_COLLISION (D2Q9: KBC-D; 3D: D3Q27 + KBC-N4) */
} OracleCase;

typedef struct OracleGrid OracleGrid;

/* LBM_initGrid (+ body labelling).  Returns NULL on allocation failure or invalid case. */
OracleGrid *luma_oracle_create(const OracleCase *c);
void        luma_oracle_destroy(OracleGrid *g);

/* nsteps calls of LBM_multi_opt.  Returns 0, or the code of the first fatal condition the
 * reference would L_ERROR on (1: BC site not within a wall, 2: pressure BC on edge/corner,
 * 3: extrapolation off grid, 4: slip site outside a wall region). */
int luma_oracle_step(OracleGrid *g, int nsteps);

/* Views of the state, in the reference's AoS layout (inc/IVector.h:94-134). */
double  *luma_oracle_f(OracleGrid *g);        /* [N*M*K*Q]  f[v + Q*(k + K*(j + M*i))] */
double  *luma_oracle_fnew(OracleGrid *g);
double  *luma_oracle_rho(OracleGrid *g);      /* [N*M*K] */
double  *luma_oracle_u(OracleGrid *g);        /* [N*M*K*dims] */
double  *luma_oracle_rho_timeav(OracleGrid *g);  /* [N*M*K]              time-averaged statistics,  */
double  *luma_oracle_ui_timeav(OracleGrid *g);   /* [N*M*K*dims]         inc/GridObj.h:98-100       */
double  *luma_oracle_uiuj_timeav(OracleGrid *g); /* [N*M*K*(3*dims-3)]                              */
int32_t *luma_oracle_lattyp(OracleGrid *g);   /* [N*M*K] eType */
int32_t *luma_oracle_wall(OracleGrid *g);     /* [N*M*K*5] {edgeCount, normalDirection, nx, ny, nz} */
double  *luma_oracle_uin(OracleGrid *g, int d); /* ux_in/uy_in/uz_in [M] */
double  *luma_oracle_pos(OracleGrid *g, int d); /* XPos/YPos/ZPos */
double   luma_oracle_omega(const OracleGrid *g);
double   luma_oracle_nu(const OracleGrid *g);
double   luma_oracle_gravity(const OracleGrid *g);
double   luma_oracle_rho_out(const OracleGrid *g);
int      luma_oracle_t(const OracleGrid *g);
void     luma_oracle_force(const OracleGrid *g, double F[3]); /* momentum-exchange force of the last step */

/* Stand-alone helpers (also used to pin host-side scalars of the product). */
double luma_oracle_velocity_ramp(const OracleCase *c, double t_dimless); /* GridUtils.cpp:1808 */
double luma_oracle_reynolds_ramp(const OracleCase *c, double t_dimless); /* GridUtils.cpp:1825 */

#ifdef __cplusplus
}
#endif
#endif
