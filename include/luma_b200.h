/* luma_b200.h -- C ABI of the B200-native replacement for LUMA's level-0 time step.
 *
 * The reference (cfdemons/LUMA v1.7.12) has no plugin or FFI interface.  The seam this library
 * plugs into is the member function
 *
 *     void GridObj::LBM_multi_opt(int subcycle = 0)        inc/GridObj.h:188
 *                                                          src/GridObj_ops_lbm_optimised.cpp:36-193
 *
 * (sole caller src/main_lbm.cpp:441) together with the halo exchange it ends with,
 *
 *     void MpiManager::mpi_communicate(int lev, int reg)   inc/MpiManager.h:238
 *                                                          src/MpiManager.cpp:631-815
 *
 * A LUMA build links a shim translation unit that is still a GridObj member (so it can read the
 * private fields f, fNew, u, rho, ux_in ... inc/GridObj.h:74-103) and forwards to the entry points
 * below; INTEGRATION.md shows that shim.  Everything here is extern "C", plain pointers and sizes,
 * int status returns (0 = ok), no exceptions cross the boundary, no torch types.
 *
 * Layouts.  Host arrays are LUMA's own: AoS, flattened as  v + Q*(k + K*(j + M*i))  for f,
 * d + D*(k + K*(j + M*i)) for u and  k + K*(j + M*i)  for rho / LatTyp (inc/IVector.h:94-134),
 * x (i) slowest.  The host keeps ownership of everything it passes; the library copies.
 * Device state (SoA populations, two lattices, packed cell words) is owned by the handle.
 *
 * Decomposition.  Level 0 is cut into x-slabs, one per process/GPU (the analogue of
 * L_MPI_XCORES = nranks, L_MPI_YCORES = L_MPI_ZCORES = 1; slab widths as
 * MpiManager::mpi_uniformDecompose, src/MpiManager.cpp:1220-1240, see luma_b200_slab()).  The
 * topology is a periodic ring in x exactly like the reference's MPI_Cart_create(periods = 1,1,1)
 * (src/MpiManager.cpp:112-136); walls/inlets/outlets in x simply make the wrap unused.
 * With nranks > 1 the per-step exchange (NCCL point-to-point) replaces mpi_communicate; the
 * host shim must not call mpi_communicate for level 0.
 */
#ifndef LUMA_B200_H
#define LUMA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LUMA_B200_ABI_VERSION 4

/* ---- status codes (luma_b200_strerror gives the text; the shim maps non-zero to L_ERROR,
 *      inc/stdafx.h:135-149) ---- */
enum {
	LUMA_B200_OK            = 0,
	LUMA_B200_EINVAL        = 1,   /* bad argument / inconsistent case */
	LUMA_B200_ECUDA         = 2,   /* CUDA runtime failure (message kept in the handle) */
	LUMA_B200_ENCCL         = 3,   /* NCCL failure */
	LUMA_B200_ENOMEM        = 4,   /* device or host allocation failed */
	LUMA_B200_EUNSUPPORTED  = 5,   /* a feature outside the level-0 path (BFL, IBM, refinement ...) */
	LUMA_B200_ESTATE        = 6,   /* call order: step before upload, comm missing with nranks > 1 ... */
	LUMA_B200_EBC_NOT_WALL  = 7,   /* a velocity/pressure site has no wall descriptor   (optimised.cpp:334-336) */
	LUMA_B200_EBC_PRESSURE_EDGE = 8, /* pressure BC on an edge/corner                   (optimised.cpp:357-359) */
	LUMA_B200_EBC_OFFGRID   = 9    /* extrapolation neighbour off grid                  (optimised.cpp:1387-1390) */
};

/* eType values the path understands (inc/Enumerations.h:84-96) */
enum { LUMA_E_SOLID = 0, LUMA_E_FLUID = 1, LUMA_E_REFINED = 2, LUMA_E_VELOCITY = 6, LUMA_E_PRESSURE = 7,
       LUMA_E_SLIP = 8, LUMA_E_EXTRAPOLATE_RIGHT = 9 };

typedef struct luma_b200 luma_b200_t;

/* Run-time image of the compile-time case (inc/definitions.h) plus the scalars LBM_initGrid
 * derived from it (src/GridObj_init_grids.cpp:336-344).  Fill with luma_b200_default_params()
 * first, then overwrite. */
typedef struct LumaCaseParams {
	uint32_t struct_size;       /* sizeof(LumaCaseParams), ABI check */
	int32_t  dims;              /* L_DIMS: 2 or 3 */
	int32_t  num_vels;          /* L_NUM_VELS: 9 (D2Q9), 19 (D3Q19) or, with kbc in 3-D, 27 (D3Q27), definitions.h:299-310 */
	int32_t  N, M, K;           /* GLOBAL level-0 size L_N, L_M, L_K (K = 1 in 2-D) */
	int32_t  rank, nranks;      /* position in the x-ring; 0,1 for the serial build */
	int32_t  x_offset, x_count; /* first owned global x-plane and number of owned planes
	                               (luma_b200_slab gives the reference's uniform split) */
	int32_t  device;            /* CUDA device ordinal for this process */
	int32_t  regularised;       /* L_REGULARISED_BOUNDARIES; 0 = forced-equilibrium velocity BC (optimised.cpp:254-270) */
	int32_t  bgksmag;           /* L_USE_BGKSMAG */
	double   csmag;             /* L_CSMAG */
	int32_t  gravity_on;        /* L_GRAVITY_ON */
	int32_t  gravity_dir;       /* L_GRAVITY_DIRECTION (0,1,2) */
	double   gravity;           /* GridObj::gravity = fd2flbm(L_GRAVITY_FORCE), inc/GridUnits.h:140 */
	double   rhoin;             /* L_RHOIN */
	double   rho_out;           /* L_RHOIN + pd2dlbm(L_PRESSURE_DELTA), optimised.cpp:343-345 */
	double   dt, dh;            /* GridObj::dt, GridObj::dh */
	double   omega;             /* GridObj::omega */
	int32_t  velocity_ramp_on;  /* L_VELOCITY_RAMP defined */
	double   velocity_ramp;     /* L_VELOCITY_RAMP */
	int32_t  reynolds_ramp_on;  /* L_REYNOLDS_RAMP defined */
	double   reynolds_ramp;     /* L_REYNOLDS_RAMP */
	double   re;                /* L_RE (only read with reynolds_ramp_on) */
	int32_t  t;                 /* GridObj::t, completed iterations at upload time */
	int32_t  time_averaged;     /* L_COMPUTE_TIME_AVERAGED_QUANTITIES (optimised.cpp:895-917) */
	int32_t  kbc;               /* L_USE_KBC_COLLISION: _LBM_kbcCollide_opt (optimised.cpp:1122-1305) replaces _LBM_collide_opt;
	                               KBC-D on D2Q9, KBC-N4 on D3Q27 (num_vels must be 27 in 3-D, and regularised 0 there:
	                               src/GridObj_init_grids.cpp:266-270) */
} LumaCaseParams;

/* Wall descriptor of one velocity/pressure/slip site, exactly what GridUtils::isWithinDomainWall
 * (src/GridUtils.cpp:1369-1430) returns for it.  `site` indexes the arrays passed to upload. */
typedef struct LumaSiteBC {
	int64_t site;
	int8_t  edge_count;         /* 1 face, 2 edge, 3 corner */
	int8_t  normal_dir;         /* eCartesianDirection of the last wall hit */
	int8_t  normal[3];          /* inward normal vector, components in {-1,0,1} */
	int8_t  pad_[3];
} LumaSiteBC;

/* Device-side construction of a case too large for (or not worth) the host object model:
 * the same labelling/initial state LBM_initGrid produces (src/GridObj_init_grids.cpp:155-384,
 * :983-1097), expressed in cell indices. */
typedef struct LumaSyntheticCase {
	int32_t wall_type[6];       /* L_WALL_LEFT, RIGHT, BOTTOM, TOP, FRONT, BACK (eType) */
	int32_t wall_cells[6];      /* L_WALL_THICKNESS_* in cells (0 = none) */
	double  u_in[3];            /* uniform inlet/lid velocity in lattice units (ud2ulbm(L_UX0) ...) */
	const double *ux_in;        /* optional profiles ux_in[j], uy_in[j], uz_in[j] of length M (e.g. the */
	const double *uy_in;        /* parabola of _LBM_initSetInletProfile, init_grids.cpp:1322-1360);    */
	const double *uz_in;        /* NULL = the uniform value u_in[d]                                    */
	int32_t no_flow;            /* L_NO_FLOW */
	int32_t has_box;            /* bounce-back body as an index box */
	int32_t box[6];             /* global i0,i1,j0,j1,k0,k1 (half-open) */
} LumaSyntheticCase;

typedef struct LumaStats {
	int64_t steps;              /* steps accepted by luma_b200_step since create */
	double  ms_last_call;       /* device time (CUDA events) of the steps submitted between the last two read points
	                               (download*, forces, stats, sync, flush): with step(n); stats() the n steps of that call */
	double  ms_per_step;        /* ms_last_call / steps in that window */
	double  mlups_last_call;    /* owned cells * steps / ms_last_call / 1e3 */
	int64_t kernel_launches;    /* kernels this library launched since create */
	int64_t halo_bytes_per_step;/* bytes this rank sends per step */
	int64_t cells;              /* owned cells */
	/* filled while profiling is on (luma_b200_set_profiling): CUDA events around every launch of the
	 * dominant kernel (k_step over the interior/all planes), accumulated since profiling was enabled */
	int64_t step_kernel_launches;
	double  step_kernel_ms;     /* summed device time of those launches */
	int64_t step_kernel_cells;  /* summed lattice updates covered by those launches */
	int64_t graph_launches;     /* CUDA-graph replays (batches of steps of launch-bound grids) since create */
} LumaStats;

#define LUMA_B200_F   1u
#define LUMA_B200_RHO 2u
#define LUMA_B200_U   4u

/* ---- life cycle.  luma_b200_create stores a handle in *h even when it FAILS (so that luma_b200_last_error can be
 *      read): the caller must luma_b200_destroy it in that case too; *h is NULL only for argument errors caught
 *      before any allocation (LUMA_B200_EINVAL from the parameter checks, LUMA_B200_ENOMEM for the handle itself). ---- */
void luma_b200_default_params(LumaCaseParams *p);
int  luma_b200_create(luma_b200_t **h, const LumaCaseParams *p);
void luma_b200_destroy(luma_b200_t *h);

/* The reference's uniform x-decomposition (MpiManager::mpi_uniformDecompose,
 * src/MpiManager.cpp:1220-1240: ceil(N/G) planes per rank, the last takes the remainder). */
int  luma_b200_slab(int32_t N, int32_t nranks, int32_t rank, int32_t *x_offset, int32_t *x_count);

/* ---- multi-GPU: attach this process to the ring.  unique_id = the 128 bytes of an ncclUniqueId
 *      created by rank 0 (luma_b200_comm_unique_id) and broadcast by the host (MPI_Bcast in a
 *      LUMA MPI build, torch.distributed in bench.py).  One unique id per handle (an id bootstraps
 *      exactly one communicator).  Replaces MpiManager::mpi_init / mpi_buffer_size for level 0. ---- */
int  luma_b200_comm_unique_id(void *unique_id_128);
int  luma_b200_comm_init(luma_b200_t *h, const void *unique_id_128);

/* Device-initiated exchange (optional, after luma_b200_comm_init and before upload/init): the ranks publish IPC
 * handles of their lattices, the host hands every rank the blobs of its left (rank-1) and right (rank+1) ring
 * neighbours (MPI_Allgather / torch.distributed.all_gather), and from then on luma_b200_step stores the outgoing
 * populations straight into the neighbour GPU's ghost planes over NVLink and signals arrival with a flag -- no
 * communication-library call per step.  Needs peer access between neighbouring GPUs (NVSwitch: always); without
 * attach the NCCL send/recv path below runs.  Results are identical either way.
 * The stores are part of the epilogue of the face-plane kernels (compute and transfer are one kernel; a one-thread launch
 * publishes the arrival flags); LUMA_B200_FUSED_HALO=0 in the environment at attach time selects the older form, a separate
 * copy kernel after the face kernels.  All three transports are parity-tested (tests/test_gpu_multi.py) and measured
 * (profiles/r02_halo_transports_n2.txt). */
#define LUMA_B200_P2P_BLOB_BYTES 256
int  luma_b200_p2p_export(luma_b200_t *h, void *blob_256);
int  luma_b200_p2p_attach(luma_b200_t *h, const void *left_blob_256, const void *right_blob_256);

/* The per-step exchange of this rank, in issue order (one NCCL group): only the populations that
 * cross a slab face travel -- c_x = +1 to the +x neighbour (D3Q19 v = 0,6,8,14,17; D2Q9 0,4,6) from
 * the last owned plane into the neighbour's low ghost plane, c_x = -1 (1,7,9,15,16; 1,5,7) from the
 * first owned plane into the neighbour's high ghost plane -- each one contiguous run of M*K doubles
 * of the SoA lattice.  The reference sends all Q populations of every halo site
 * (src/Mpi_buffer_pack.cpp:72-96).  `plane` is a local plane index (0 = low ghost, 1..x_count owned,
 * x_count+1 = high ghost).  Host-only; needs no device. */
typedef struct LumaHaloMsg {
	int32_t is_send;            /* 1 send, 0 receive */
	int32_t peer;               /* rank of the other side */
	int32_t pop;                /* population index v */
	int32_t plane;              /* local plane read (send) or written (receive) */
} LumaHaloMsg;
int  luma_b200_halo_plan(const LumaCaseParams *p, LumaHaloMsg *msgs, int32_t capacity, int32_t *count);

/* ---- state in: everything LBM_multi_opt reads (GridObj fields, inc/GridObj.h:74-125).
 *      Arrays cover this rank's owned planes, preceded/followed by `halo` extra x-planes
 *      (0 for the serial build, 1 for LUMA's MPI build whose local arrays carry recv layers,
 *      src/MpiManager.cpp:277-282).  u_aos/rho give the stored macroscopic fields (they matter for
 *      sites the kernel never updates).  ux_in/uy_in/uz_in have M entries (NULL = zeros).
 *      f_aos == NULL declares f = feq(rho, u) at EVERY site -- the state LBM_initGrid leaves at t = 0
 *      (src/GridObj_init_grids.cpp:310-333) when no site had its u changed afterwards (L_NO_FLOW builds; bodies
 *      labelled after initialisation zero the u of their sites, src/ObjectManager.cpp:333-338): the device then
 *      evaluates _LBM_equilibrium_opt itself, bit for bit, and 8*Q bytes per site never cross the PCIe bus.
 *      Steps accepted by luma_b200_step but not yet submitted are dropped (the state is replaced). ---- */
int  luma_b200_upload(luma_b200_t *h, int32_t halo,
                      const double *f_aos, const double *rho, const double *u_aos,
                      const int32_t *lattyp,
                      const LumaSiteBC *bc_sites, size_t n_bc,
                      const double *ux_in, const double *uy_in, const double *uz_in);

/* ---- state built on the device (benchmark shapes; also usable as a fast LBM_initGrid) ---- */
int  luma_b200_init_synthetic(luma_b200_t *h, const LumaSyntheticCase *c);

/* ---- nsteps calls of LBM_multi_opt (+ the exchange).  NEVER waits for the GPU: the call advances GridObj::t /
 *      omega / nu on the host (luma_b200_get_time) and queues the steps.  Submission is lazy -- the most recent step
 *      is held back until the next call or read point, because the last step before the host looks at the fields is
 *      the one that stores rho,u (all others keep them in registers; the host may look at main_lbm.cpp:449-561), and
 *      launch-bound grids collect LUMA_B200_GRAPH_STEPS steps into one CUDA-graph launch even when the host calls
 *      once per step (src/main_lbm.cpp:441) -- single rank, and slabs with the device-initiated exchange.  Every entry point that reads state -- download*, forces, stats, sync --
 *      submits what is held back first; luma_b200_flush submits without reading or waiting (call it before a long
 *      stretch of host work).  A halo time-out (dead ring neighbour) is reported by the next call that notices it.
 *      Several ranks: a rank's step t+1 needs its ring neighbours' step t, so a rank must not block on another rank (MPI_Barrier,
 *      MPI_Recv ...) while it still holds an accepted step back that the other rank's pending read needs.  Hosts that take
 *      their read points collectively (LUMA's loop does: every rank steps and writes output at the same t) never can; any
 *      other host calls luma_b200_flush before it blocks on a peer. ---- */
int  luma_b200_step(luma_b200_t *h, int32_t nsteps);
int  luma_b200_flush(luma_b200_t *h);

/* ---- state out, same layout/halo convention as upload; only owned planes are written.
 *      `what` = LUMA_B200_F | LUMA_B200_RHO | LUMA_B200_U; unused pointers may be NULL. ---- */
int  luma_b200_download(luma_b200_t *h, int32_t halo, unsigned what,
                        double *f_aos, double *rho, double *u_aos);
int  luma_b200_download_lattyp(luma_b200_t *h, int32_t halo, int32_t *lattyp);
/* Asynchronous download for hosts that write output while the next steps run (the IO points of
 * src/main_lbm.cpp:449-561 moved off the critical path): the fields are snapshotted on the device in the host
 * layout, in stream order after the steps issued so far, and copied on a separate stream; the arrays
 * (pinned host memory for a truly asynchronous copy) may be read after luma_b200_download_wait().  A second
 * download_async before the wait queues behind the first. */
int  luma_b200_download_async(luma_b200_t *h, int32_t halo, unsigned what,
                              double *f_aos, double *rho, double *u_aos);
int  luma_b200_download_wait(luma_b200_t *h);

/* ---- time-averaged statistics (handles created with time_averaged = 1): rho_timeav [cells],
 *      ui_timeav [cells*D], uiuj_timeav [cells*(3D-3)], the reference's arrays (inc/GridObj.h:93-95,
 *      written by io_hdf5 / io_lite).  They start at zero (init_grids.cpp:304-306); upload_timeav lets a
 *      host that kept them across a restart hand them back.  NULL pointers are skipped. ---- */
int  luma_b200_download_timeav(luma_b200_t *h, int32_t halo, double *rho_timeav, double *ui_timeav, double *uiuj_timeav);
int  luma_b200_upload_timeav(luma_b200_t *h, int32_t halo, const double *rho_timeav, const double *ui_timeav, const double *uiuj_timeav);

/* ---- binary restart: t, omega, nu, rho, u, f and -- for handles created with time_averaged -- the three
 *      averages of this rank's owned planes, raw little-endian doubles in the reference's array layouts behind a 64-byte
 *      header; the content of GridObj::io_restart (src/GridObj_ops_io.cpp:406-640), which spends 17 decimal digits of ASCII
 *      per value.  One file per rank.  restart_read needs the geometry first (luma_b200_upload of the freshly initialised
 *      grid with f_aos = NULL, or luma_b200_init_synthetic) and replaces t and the fields; the run continues bit-identically. ---- */
int  luma_b200_restart_write(luma_b200_t *h, const char *path);
int  luma_b200_restart_read(luma_b200_t *h, const char *path);

/* ---- scalars the host object keeps in step with the device (GridObj::t, ::omega, ::nu) ---- */
int  luma_b200_get_time(luma_b200_t *h, int32_t *t, double *omega, double *nu);

/* ---- momentum-exchange force on bounce-back bodies accumulated by the LAST step
 *      (ObjectManager::computeLiftDrag(i,j,k,g), src/ObjectManager.cpp:93-164), this rank's part: the links whose
 *      FLUID end this rank owns (the sum over ranks is the reference's sum).  Needs no barrier between the ranks:
 *      only this rank's own planes of the previous lattice are read. ---- */
int  luma_b200_forces(luma_b200_t *h, double F[3]);

int  luma_b200_stats(luma_b200_t *h, LumaStats *s);
/* on != 0: bracket each launch of the dominant kernel with CUDA events on its stream (read back by
 * luma_b200_stats); enabling resets the accumulators.  Off by default. */
int  luma_b200_set_profiling(luma_b200_t *h, int32_t on);
int  luma_b200_sync(luma_b200_t *h);
/* device self-test of the constant-divisor division used for x/cs^2 and x/(2cs^4) (lattice.cuh,
 * tests/test_constdiv_exact.py): compares it with IEEE `/` on n pseudo-random operands. */
int  luma_b200_selftest_div_const(int32_t device, int64_t n, uint64_t seed, int64_t *mismatches);
const char *luma_b200_strerror(int code);
const char *luma_b200_last_error(luma_b200_t *h);   /* detail of the last non-zero return */
int  luma_b200_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* LUMA_B200_H */
