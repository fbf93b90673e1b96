/* GridObj_ops_lbm_b200.cpp -- the drop-in: GridObj::LBM_multi_opt forwarding to libluma_b200.so.
 *
 * Build it INTO LUMA (it includes LUMA's own headers and is a GridObj member, so it can read the
 * private fields f, u, rho, ux_in ..., inc/GridObj.h:74-103) and compile LUMA's
 * src/GridObj_ops_lbm_optimised.cpp with  -DLBM_multi_opt=LBM_multi_opt_cpu  so that its CPU body
 * (src/GridObj_ops_lbm_optimised.cpp:36-193) keeps a different name while every other function of that
 * file (_LBM_equilibrium_opt is needed by LBM_initGrid and io_restart) stays linked.  Nothing else
 * in LUMA changes: inc/definitions.h still fixes the case at compile time, GridObj/GridManager/
 * ObjectManager build the grid, label walls and bodies, and the IO code reads the same host arrays.
 *
 *   first call : LumaCaseParams from the macros + members, wall descriptors of the velocity/pressure/slip
 *                sites from GridUtils::isWithinDomainWall, luma_b200_upload of rho, u, LatTyp -- and of f
 *                unless the grid is freshly initialised (t == 0 of an L_NO_FLOW build: f = feq(rho,u) at every
 *                site, src/GridObj_init_grids.cpp:310-333, which the device evaluates itself bit for bit).
 *   every call : luma_b200_step(1) -- it only queues the step and NEVER waits for the GPU, so LUMA's loop
 *                (src/main_lbm.cpp:422-572) runs ahead of the device and launch-bound grids are replayed as
 *                CUDA-graph batches although the steps arrive one call at a time; GridObj::t, omega, nu
 *                follow the library's host-side values.
 *   host sync  : rho, u (and f when a restart file is due) are downloaded into the GridObj arrays
 *                whenever main() is about to read them -- t % L_GRID_OUT_FREQ, L_PROBE_OUT_FREQ,
 *                L_EXTRA_OUT_FREQ, L_RESTART_OUT_FREQ (src/main_lbm.cpp:449-561) -- or on demand with
 *                LBM_multi_opt(LUMA_B200_SYNC_HOST).  These are the only points where the host waits.
 *   lift/drag  : with L_LD_OUT the momentum-exchange force of the step (ObjectManager::computeLiftDrag,
 *                src/ObjectManager.cpp:93-164) is fetched with luma_b200_forces at the steps
 *                io_writeForcesOnObjects writes it (t % L_EXTRA_OUT_FREQ, src/main_lbm.cpp:505-517) and stored in
 *                ObjectManager::bbbForceOnObjectX/Y/Z (GridObj is a friend of ObjectManager, inc/ObjectManager.h:44).
 *
 * Serial build (L_BUILD_FOR_MPI undefined): one process, one GPU.  MPI build: one rank per GPU,
 * L_MPI_XCORES = ranks, L_MPI_YCORES = L_MPI_ZCORES = 1; the ncclUniqueId is broadcast with MPI_Bcast
 * and mpi_communicate is NOT called for level 0 (the library exchanges the slab faces itself).
 */
#include "LUMA/inc/stdafx.h"
#include "LUMA/inc/GridObj.h"
#include "LUMA/inc/GridUtils.h"
#include "LUMA/inc/ObjectManager.h"
#ifdef L_BUILD_FOR_MPI
#include "LUMA/inc/MpiManager.h"
#endif

#include "luma_b200.h"

#include <cstdlib>
#include <vector>

#ifndef LUMA_B200_SYNC_HOST
#define LUMA_B200_SYNC_HOST (-1)      /* LBM_multi_opt(LUMA_B200_SYNC_HOST): download f, rho, u; no time step */
#define LUMA_B200_SYNC_MACRO (-2)     /* LBM_multi_opt(LUMA_B200_SYNC_MACRO): download rho, u only; no time step */
#endif

static_assert(sizeof(eType) == sizeof(int32_t), "LatTyp is handed over as int32");

#include <chrono>

namespace
{
	luma_b200_t *g_dev = nullptr;
	double g_create_seconds = 0.0, g_upload_seconds = 0.0;
	double now_seconds() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

	void check(int rc, const char *what)
	{
		if (rc == LUMA_B200_OK) return;
		std::string msg = std::string("luma_b200 ") + what + ": " + luma_b200_strerror(rc);
		if (g_dev) msg += std::string(" [") + luma_b200_last_error(g_dev) + "]";
		L_ERROR(msg, GridUtils::logfile);      /* log, MPI_Finalize, exit -- inc/stdafx.h:135-149 */
	}
}

/* for hosts that want the library's own statistics or an explicit luma_b200_sync / luma_b200_flush: the handle the
 * shim created, and how long its one-off calls took (luma_b200_create incl. CUDA context start-up; luma_b200_upload) */
extern "C" luma_b200_t *luma_b200_shim_handle(double *create_seconds, double *upload_seconds)
{
	if (create_seconds) *create_seconds = g_create_seconds;
	if (upload_seconds) *upload_seconds = g_upload_seconds;
	return g_dev;
}

void GridObj::LBM_multi_opt(int subcycle)
{
	if (level != 0)
		L_ERROR("luma_b200 accelerates level 0 only (L_NUM_LEVELS must be 0)", GridUtils::logfile);

	const int halo =
#ifdef L_BUILD_FOR_MPI
		1;      /* local arrays carry one recv layer each side in x, src/MpiManager.cpp:277-282 */
#else
		0;
#endif

	if (!g_dev)
	{
		LumaCaseParams p;
		luma_b200_default_params(&p);
		p.dims = L_DIMS;
		p.num_vels = L_NUM_VELS;
		p.N = L_N; p.M = L_M; p.K = L_K;
		int local_rank = 0;
#ifdef L_BUILD_FOR_MPI
		MpiManager *mpim = MpiManager::getInstance();
		if (L_MPI_YCORES != 1 || L_MPI_ZCORES != 1)
			L_ERROR("luma_b200 decomposes along x only: set L_MPI_YCORES = L_MPI_ZCORES = 1", GridUtils::logfile);
		p.rank = mpim->my_rank; p.nranks = mpim->num_ranks;
		{
			/* the GPU of a rank is chosen by its rank WITHIN the node */
			MPI_Comm node;
			MPI_Comm_split_type(mpim->world_comm, MPI_COMM_TYPE_SHARED, p.rank, MPI_INFO_NULL, &node);
			MPI_Comm_rank(node, &local_rank);
			MPI_Comm_free(&node);
		}
#endif
		check(luma_b200_slab(p.N, p.nranks, p.rank, &p.x_offset, &p.x_count), "slab");
		if (p.x_count + 2 * halo != N_lim)
			L_ERROR("luma_b200: slab width differs from the host decomposition (use the uniform decomposition)", GridUtils::logfile);
		const char *dev = getenv("LUMA_B200_DEVICE");
		const char *ndev = getenv("LUMA_B200_DEVICES_PER_NODE");
		p.device = dev ? atoi(dev) : (ndev && atoi(ndev) > 0 ? local_rank % atoi(ndev) : local_rank);
#ifdef L_REGULARISED_BOUNDARIES
		p.regularised = 1;
#else
		p.regularised = 0;
#endif
#ifdef L_USE_BGKSMAG
		p.bgksmag = 1;
#endif
		p.csmag = L_CSMAG;
#ifdef L_GRAVITY_ON
		p.gravity_on = 1;
#endif
		p.gravity_dir = static_cast<int>(L_GRAVITY_DIRECTION);
		p.gravity = gravity;
		p.rhoin = L_RHOIN;
		p.rho_out = L_RHOIN;
#ifdef L_PRESSURE_DELTA
		p.rho_out += GridUnits::pd2dlbm(L_PRESSURE_DELTA, this);      /* optimised.cpp:343-345 */
#endif
		p.dt = dt; p.dh = dh;
		p.omega = omega;
#ifdef L_VELOCITY_RAMP
		p.velocity_ramp_on = 1; p.velocity_ramp = L_VELOCITY_RAMP;
#endif
#ifdef L_REYNOLDS_RAMP
		p.reynolds_ramp_on = 1; p.reynolds_ramp = L_REYNOLDS_RAMP;
#endif
		p.re = static_cast<double>(L_RE);
		p.t = t;
#ifdef L_COMPUTE_TIME_AVERAGED_QUANTITIES
		p.time_averaged = 1;
#endif
#ifdef L_USE_KBC_COLLISION
		p.kbc = 1;      /* _LBM_kbcCollide_opt instead of _LBM_collide_opt; L_NUM_VELS is 27 in 3-D */
#endif
#if defined(L_IBM_ON) || (L_NUM_LEVELS != 0)
		L_ERROR("luma_b200: IBM and grid refinement are outside the accelerated path", GridUtils::logfile);
#endif
		const double t_create = now_seconds();
		check(luma_b200_create(&g_dev, &p), "create");
		g_create_seconds = now_seconds() - t_create;
#ifdef L_BUILD_FOR_MPI
		{
			char id[128];
			if (p.rank == 0) check(luma_b200_comm_unique_id(id), "comm_unique_id");
			MPI_Bcast(id, 128, MPI_CHAR, 0, mpim->world_comm);
			check(luma_b200_comm_init(g_dev, id), "comm_init");
			/* Device-initiated halo exchange (CUDA IPC peer stores: ranks of ONE node): every rank publishes its IPC blob
			 * and takes its ring neighbours'.  The choice of transport is collective -- all ranks attach or none does;
			 * when any rank cannot (several nodes, no peer access) the NCCL send/recv exchange stays in place. */
			std::vector<char> mine(LUMA_B200_P2P_BLOB_BYTES), all((size_t)LUMA_B200_P2P_BLOB_BYTES * p.nranks);
			int can = getenv("LUMA_B200_HALO_NCCL") ? 0 : 1, all_can = 0;
			if (can && luma_b200_p2p_export(g_dev, &mine[0]) != LUMA_B200_OK) can = 0;
			{
				/* one node only: the node-local communicator must be the whole world */
				MPI_Comm node; int nsize = 0;
				MPI_Comm_split_type(mpim->world_comm, MPI_COMM_TYPE_SHARED, p.rank, MPI_INFO_NULL, &node);
				MPI_Comm_size(node, &nsize);
				MPI_Comm_free(&node);
				if (nsize != p.nranks) can = 0;
			}
			MPI_Allreduce(&can, &all_can, 1, MPI_INT, MPI_MIN, mpim->world_comm);
			if (all_can)
			{
				MPI_Allgather(&mine[0], LUMA_B200_P2P_BLOB_BYTES, MPI_CHAR, &all[0], LUMA_B200_P2P_BLOB_BYTES, MPI_CHAR, mpim->world_comm);
				const int left = (p.rank - 1 + p.nranks) % p.nranks, right = (p.rank + 1) % p.nranks;
				int ok = luma_b200_p2p_attach(g_dev, &all[(size_t)left * LUMA_B200_P2P_BLOB_BYTES], &all[(size_t)right * LUMA_B200_P2P_BLOB_BYTES]) == LUMA_B200_OK;
				int all_ok = 0, any_ok = 0;
				MPI_Allreduce(&ok, &all_ok, 1, MPI_INT, MPI_MIN, mpim->world_comm);
				MPI_Allreduce(&ok, &any_ok, 1, MPI_INT, MPI_MAX, mpim->world_comm);
				if (any_ok && !all_ok)      /* a ring with mixed transports would dead-lock: give up cleanly */
					L_ERROR("luma_b200: peer mappings could be opened on some ranks only; rerun with LUMA_B200_HALO_NCCL=1", GridUtils::logfile);
			}
		}
#endif
		/* wall descriptors exactly as _LBM_regularised_opt would obtain them (optimised.cpp:334) */
		std::vector<LumaSiteBC> bc;
		std::vector<int> nv(3, 0);
		for (int i = 0; i < N_lim; ++i) for (int j = 0; j < M_lim; ++j) for (int k = 0; k < K_lim; ++k)
		{
			const int64_t id = k + (int64_t)j * K_lim + (int64_t)i * K_lim * M_lim;
			const eType ty = LatTyp[id];
			if (ty != eVelocity && ty != ePressure && ty != eSlip) continue;
			eCartesianDirection nd = eXDirection; unsigned int ec = 0;
			nv[0] = nv[1] = nv[2] = 0;
			LumaSiteBC s = { id, 0, 0, { 0, 0, 0 }, { 0, 0, 0 } };
			if (GridUtils::isWithinDomainWall(XPos[i], YPos[j], ZPos[k], &nv, &nd, &ec))
			{
				s.edge_count = (int8_t)ec; s.normal_dir = (int8_t)nd;
				s.normal[0] = (int8_t)nv[0]; s.normal[1] = (int8_t)nv[1]; s.normal[2] = (int8_t)nv[2];
			}
			bc.push_back(s);
		}
		/* A freshly initialised grid holds f = feq(rho,u) everywhere (LBM_initGrid, init_grids.cpp:310-333) provided no
		 * site had its u changed after that loop: true for L_NO_FLOW builds (u = 0 except on the velocity walls, and
		 * bodies are labelled on fluid sites whose u is already 0, src/ObjectManager.cpp:333-338).  Then f is not
		 * uploaded: the device evaluates the same expression.  After a restart, or without L_NO_FLOW, f travels. */
		const double *f_host = &f[0];
#if defined(L_NO_FLOW) && !defined(L_INIT_VELOCITY_FROM_FILE)
		if (t == 0 && !getenv("LUMA_B200_UPLOAD_F")) f_host = nullptr;
#endif
		const double t_upload = now_seconds();
		check(luma_b200_upload(g_dev, halo, f_host, &rho[0], &u[0], reinterpret_cast<const int32_t *>(&LatTyp[0]),
			bc.empty() ? nullptr : &bc[0], bc.size(), &ux_in[0], &uy_in[0], &uz_in[0]), "upload");
		g_upload_seconds = now_seconds() - t_upload;
	}

	if (subcycle == LUMA_B200_SYNC_HOST)
	{
		check(luma_b200_download(g_dev, halo, LUMA_B200_F | LUMA_B200_RHO | LUMA_B200_U, &f[0], &rho[0], &u[0]), "download");
#ifdef L_COMPUTE_TIME_AVERAGED_QUANTITIES
		check(luma_b200_download_timeav(g_dev, halo, &rho_timeav[0], &ui_timeav[0], &uiuj_timeav[0]), "download_timeav");
#endif
		return;
	}
	if (subcycle == LUMA_B200_SYNC_MACRO)
	{
		check(luma_b200_download(g_dev, halo, LUMA_B200_RHO | LUMA_B200_U, nullptr, &rho[0], &u[0]), "download");
		return;
	}

	clock_t t_start = clock();
	check(luma_b200_step(g_dev, 1), "step");                             /* queues the step; does not wait for the GPU */
	check(luma_b200_get_time(g_dev, &t, &omega, &nu), "get_time");      /* ++t, _LBM_updateReynolds (optimised.cpp:39-42,:170) */

	/* host arrays are refreshed exactly when main() reads them (src/main_lbm.cpp:449-561) */
	unsigned what = 0;
	if (t % L_GRID_OUT_FREQ == 0 || t % L_PROBE_OUT_FREQ == 0 || t % L_EXTRA_OUT_FREQ == 0) what |= LUMA_B200_RHO | LUMA_B200_U;
	if (t % L_RESTART_OUT_FREQ == 0) what |= LUMA_B200_F | LUMA_B200_RHO | LUMA_B200_U;
#ifdef L_LD_OUT
	/* the momentum-exchange force of this step, where io_writeForcesOnObjects is about to write it
	 * (src/main_lbm.cpp:505-517, src/ObjectManager_ops_io.cpp:1061-1098); with several ranks each writes its own share to
	 * its own file, exactly as the reference does */
	if (t % L_EXTRA_OUT_FREQ == 0)
	{
		double F[3] = { 0.0, 0.0, 0.0 };
		check(luma_b200_forces(g_dev, F), "forces");
		ObjectManager *objman = ObjectManager::getInstance();
		objman->bbbForceOnObjectX = F[0];
		objman->bbbForceOnObjectY = F[1];
		objman->bbbForceOnObjectZ = F[2];
	}
#endif
	if (what) check(luma_b200_download(g_dev, halo, what, &f[0], &rho[0], &u[0]), "download");
#ifdef L_COMPUTE_TIME_AVERAGED_QUANTITIES
	/* io_hdf5 / io_lite write the averages with the grid output (src/GridObj_ops_io.cpp:764-791, :1144-1262) */
	if (t % L_GRID_OUT_FREQ == 0)
		check(luma_b200_download_timeav(g_dev, halo, &rho_timeav[0], &ui_timeav[0], &uiuj_timeav[0]), "download_timeav");
#endif

	/* the reference's running average of the step time (optimised.cpp:172-183): here the host time of the call */
	const double secs = static_cast<double>(clock() - t_start) / CLOCKS_PER_SEC;
	timeav_timestep *= (t - 1);
	timeav_timestep += secs;
	timeav_timestep /= t;
}
