/* TEST INFRASTRUCTURE ONLY (oracle build) -- never part of the product path.
 *
 * LUMA fixes a case at compile time through the macros of inc/definitions.h
 * (/root/reference/LUMA/inc/definitions.h:23-352).  The reference sources are
 * compiled where they lie, read-only, so instead of editing that file the
 * oracle build force-includes ONE case header (oracle/cases/<case>.h, via
 * `g++ -include`), which sets the macros that differ from the defaults below
 * and then includes this tail.  The tail claims definitions.h's include guard
 * (LBM_DEFINITIONS_H, definitions.h:28-29), so the stock case in the reference
 * tree is skipped, fills in a default for every macro the sources consult, and
 * resolves the dependent options the same way definitions.h:299-350 does.
 *
 * Switch-type macros (L_GRAVITY_ON, L_USE_BGKSMAG, L_REGULARISED_BOUNDARIES,
 * L_NO_FLOW, L_VELOCITY_RAMP, L_REYNOLDS_RAMP, L_ENABLE_OPENMP, L_LD_OUT, ...)
 * are simply defined, or not, by the case header.  The build never defines
 * L_BUILD_FOR_MPI, L_IBM_ON, L_HDF5_OUTPUT or L_GEOMETRY_FILE: those
 * subsystems are outside the hot path.  L_USE_KBC_COLLISION is a
 * case switch (last row).
 */
#ifndef LBM_DEFINITIONS_H
#define LBM_DEFINITIONS_H

#include <time.h>
#include <iostream>
#include <fstream>
#include <vector>
#include <iomanip>
#include <math.h>
#include <string>
#include <limits>
#include <mpi.h>

/* ---- derived grid sizes (definitions.h:46-50) ---- */
#define L_N static_cast<int>((L_BX) * L_RESOLUTION)
#define L_M static_cast<int>((L_BY) * L_RESOLUTION)
#define L_K static_cast<int>((L_BZ) * L_RESOLUTION)
#define L_COARSE_SITE_WIDTH (1.0 / static_cast<double>(L_RESOLUTION))

/* ---- output cadence: the oracle driver does its own dumps ---- */
#ifndef L_GRID_OUT_FREQ
#define L_GRID_OUT_FREQ 1000000000
#endif
#ifndef L_EXTRA_OUT_FREQ
#define L_EXTRA_OUT_FREQ 1000000000
#endif
#define L_OUTPUT_PRECISION 17
#define L_RESTART_OUT_FREQ 1000000000
#define L_PROBE_OUT_FREQ 1000000000
#define L_PROBE_NUM_X 0
#define L_PROBE_NUM_Y 0
#define L_PROBE_NUM_Z 0
#define L_PROBE_MIN_X 0.0
#define L_PROBE_MIN_Y 0.0
#define L_PROBE_MIN_Z 0.0
#define L_PROBE_MAX_X 0.0
#define L_PROBE_MAX_Y 0.0
#define L_PROBE_MAX_Z 0.0

/* ---- forcing ---- */
#ifndef L_GRAVITY_FORCE
#define L_GRAVITY_FORCE 0.0
#endif
#ifndef L_GRAVITY_DIRECTION
#define L_GRAVITY_DIRECTION eXDirection
#endif

/* ---- collision ---- */
#ifndef L_CSMAG
#define L_CSMAG 0.3
#endif

/* ---- time ---- */
#ifndef L_TOTAL_TIMESTEPS
#define L_TOTAL_TIMESTEPS 100
#endif

/* ---- MPI layout (unused: serial build) ---- */
#define L_MPI_XCORES 1
#define L_MPI_YCORES 1
#define L_MPI_ZCORES 1
#define L_MPI_SD_MAX_ITER 1
#define L_MPI_TOP_XCORES 1
#define L_MPI_TOP_YCORES 1
#define L_MPI_TOP_ZCORES 1

/* ---- lattice / domain ---- */
#ifndef L_BZ
#define L_BZ 1.0
#endif
#define L_PHYSICAL_U 1.0
#define L_PHYSICAL_RHO 1000.0

/* ---- fluid ---- */
#ifndef L_UX0
#define L_UX0 1.0
#endif
#ifndef L_UY0
#define L_UY0 0.0
#endif
#ifndef L_UZ0
#define L_UZ0 0.0
#endif
#ifndef L_RHOIN
#define L_RHOIN 1
#endif
#if !defined(L_RE) && !defined(L_NU)
#error "case header must define L_RE or L_NU"
#endif
#ifndef L_RE
#define L_RE 1
#endif

/* ---- FEM constants referenced by off-path TUs ---- */
#define L_NB_ALPHA 0.25
#define L_NB_DELTA 0.5
#define L_RELAX 0.5

/* ---- walls ---- */
#ifndef L_PRESSURE_DELTA
#define L_PRESSURE_DELTA 0.0
#endif
#ifndef L_WALL_THICKNESS_BOTTOM
#define L_WALL_THICKNESS_BOTTOM (1.0 * L_COARSE_SITE_WIDTH)
#endif
#ifndef L_WALL_THICKNESS_TOP
#define L_WALL_THICKNESS_TOP (1.0 * L_COARSE_SITE_WIDTH)
#endif
#ifndef L_WALL_THICKNESS_LEFT
#define L_WALL_THICKNESS_LEFT (1.0 * L_COARSE_SITE_WIDTH)
#endif
#ifndef L_WALL_THICKNESS_RIGHT
#define L_WALL_THICKNESS_RIGHT (1.0 * L_COARSE_SITE_WIDTH)
#endif
#ifndef L_WALL_THICKNESS_FRONT
#define L_WALL_THICKNESS_FRONT (1.0 * L_COARSE_SITE_WIDTH)
#endif
#ifndef L_WALL_THICKNESS_BACK
#define L_WALL_THICKNESS_BACK (1.0 * L_COARSE_SITE_WIDTH)
#endif

/* ---- no refinement: level 0 only ---- */
#define L_NUM_LEVELS 0
#define L_NUM_REGIONS 1
#define L_PADDING_X_MIN 0.0
#define L_PADDING_X_MAX 0.0
#define L_PADDING_Y_MIN 0.0
#define L_PADDING_Y_MAX 0.0
#define L_PADDING_Z_MIN 0.0
#define L_PADDING_Z_MAX 0.0
static double cRefStartX[1][1] = { { 0.0 } };
static double cRefEndX[1][1] = { { 0.0 } };
static double cRefStartY[1][1] = { { 0.0 } };
static double cRefEndY[1][1] = { { 0.0 } };
static double cRefStartZ[1][1] = { { 0.0 } };
static double cRefEndZ[1][1] = { { 0.0 } };

/* ---- probes (off path) ---- */
const static int cNumProbes[3] = { L_PROBE_NUM_X, L_PROBE_NUM_Y, L_PROBE_NUM_Z };
const static double cProbeLimsX[2] = { L_PROBE_MIN_X, L_PROBE_MAX_X };
const static double cProbeLimsY[2] = { L_PROBE_MIN_Y, L_PROBE_MAX_Y };
const static double cProbeLimsZ[2] = { L_PROBE_MIN_Z, L_PROBE_MAX_Z };

/* ---- dependent options (definitions.h:299-337) ---- */
#if (L_DIMS == 3)
#ifdef L_USE_KBC_COLLISION
#define L_NUM_VELS 27
#else
#define L_NUM_VELS 19
#endif
#define L_MPI_DIRS 26
#else
#define L_NUM_VELS 9
#define L_MPI_DIRS 8
#undef L_BZ
#define L_BZ 0
#undef L_K
#define L_K 1
#define L_BLOCK_MIN_Z 0.0
#define L_BLOCK_MAX_Z 0.0
#undef L_UZ0
#define L_UZ0 0.0
#endif

#endif /* LBM_DEFINITIONS_H */
