/* TEST INFRASTRUCTURE ONLY (oracle build) -- never part of the product path.
 *
 * Driver that replaces src/main_lbm.cpp and src/GridObj_ops_io.cpp when the
 * UNMODIFIED LUMA reference sources are compiled, where they lie under
 * /root/reference/LUMA, into oracle/_ref/luma_ref_<case> (recipe: oracle/Makefile).
 * It runs the reference's own level-0 time step, GridObj::LBM_multi_opt
 * (src/GridObj_ops_lbm_optimised.cpp:36), and dumps raw little-endian state so the
 * C restatement (oracle/luma_oracle.c) and the CUDA path can be compared with it.
 *
 * Init sequence mirrors src/main_lbm.cpp:58-346 for a serial, level-0-only run:
 * GridManager::getInstance() -> new GridObj(0) -> setGridHierarchy ->
 * ObjectManager::getInstance(Grids) -> (body labelling) -> time loop :422-572.
 *
 * With -DLUMA_DROPIN the same driver is linked against luma_b200/host/GridObj_ops_lbm_b200.cpp
 * instead of the reference's CPU LBM_multi_opt (oracle/Makefile target `dropin`): the unmodified LUMA
 * host then steps on the GPU -- the drop-in demonstration checked by tests/test_gpu_dropin.py.
 *
 * usage:  luma_ref_<case> dump  <outdir> <step[,step...]>
 *         luma_ref_<case> bench <warmup> <steps>
 */
#include "LUMA/inc/stdafx.h"
#include "LUMA/inc/GridObj.h"
#include "LUMA/inc/ObjectManager.h"
#include "LUMA/inc/PCpts.h"

#ifdef LUMA_DROPIN
#include "luma_b200.h"
extern "C" luma_b200_t *luma_b200_shim_handle(double *create_seconds, double *upload_seconds);   /* luma_b200/host/GridObj_ops_lbm_b200.cpp */
#endif
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sstream>
#include <sys/stat.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ---- symbols normally provided by the two excluded TUs ---- */
std::string GridUtils::path_str;                       /* src/main_lbm.cpp:55 */
extern "C" void dgetrf_(int *, int *, double *, int *, int *, int *) {}
extern "C" void dgetrs_(char *, int *, int *, double *, int *, int *, double *, int *, int *) {}
void GridObj::io_textout(std::string) {}
void GridObj::io_fgaout() {}
void GridObj::_io_fgaout(int) {}
void GridObj::io_restart(eIOFlag) {}
void GridObj::io_probeOutput() {}
int GridObj::io_hdf5(double) { return 0; }

/* ---- read-only access to ObjectManager's private momentum-exchange accumulators
 *      (inc/ObjectManager.h:97-99) through explicit-instantiation access ---- */
template <typename Tag, typename Tag::type Member>
struct LumaOraclePeek { friend typename Tag::type luma_oracle_peek(Tag) { return Member; } };
struct PeekFx { typedef double ObjectManager::*type; friend type luma_oracle_peek(PeekFx); };
struct PeekFy { typedef double ObjectManager::*type; friend type luma_oracle_peek(PeekFy); };
struct PeekFz { typedef double ObjectManager::*type; friend type luma_oracle_peek(PeekFz); };
template struct LumaOraclePeek<PeekFx, &ObjectManager::bbbForceOnObjectX>;
template struct LumaOraclePeek<PeekFy, &ObjectManager::bbbForceOnObjectY>;
template struct LumaOraclePeek<PeekFz, &ObjectManager::bbbForceOnObjectZ>;

static void write_raw(const std::string &path, const void *p, size_t bytes)
{
	FILE *fh = fopen(path.c_str(), "wb");
	if (!fh || fwrite(p, 1, bytes, fh) != bytes) { perror(path.c_str()); exit(2); }
	fclose(fh);
}

/* GridObj::io_lite is a member, so it can read the private fields (inc/GridObj.h:52-103).
 * The driver uses it as the dump hook: Tag is "<outdir>/<prefix>", tval>=0 selects what to write. */
void GridObj::io_lite(double tval, std::string Tag)
{
	const size_t nsites = (size_t)N_lim * M_lim * K_lim;
	if (tval < 0.0)
	{
		/* one-off description of the initial state */
		std::vector<int> lt(nsites);
		for (size_t s = 0; s < nsites; ++s) lt[s] = (int)LatTyp[s];
		write_raw(Tag + ".lattyp.i32", lt.data(), nsites * sizeof(int));
		write_raw(Tag + ".xpos.f64", XPos.data(), XPos.size() * sizeof(double));
		write_raw(Tag + ".ypos.f64", YPos.data(), YPos.size() * sizeof(double));
		write_raw(Tag + ".zpos.f64", ZPos.data(), ZPos.size() * sizeof(double));
		write_raw(Tag + ".ux_in.f64", ux_in.data(), ux_in.size() * sizeof(double));
		write_raw(Tag + ".uy_in.f64", uy_in.data(), uy_in.size() * sizeof(double));
		write_raw(Tag + ".uz_in.f64", uz_in.data(), uz_in.size() * sizeof(double));

		/* per-site wall descriptors exactly as _LBM_regularised_opt sees them
		 * (GridUtils::isWithinDomainWall, src/GridUtils.cpp:1369): 5 ints per site
		 * {edgeCount, normalDirection, nx, ny, nz}, zeros where not within a wall */
		std::vector<int> bc(nsites * 5, 0);
		std::vector<int> nv(3, 0);
		for (int i = 0; i < N_lim; ++i) for (int j = 0; j < M_lim; ++j) for (int k = 0; k < K_lim; ++k)
		{
			eCartesianDirection nd; unsigned int ec;
			size_t id = (size_t)k + (size_t)j * K_lim + (size_t)i * K_lim * M_lim;
			if (GridUtils::isWithinDomainWall(XPos[i], YPos[j], ZPos[k], &nv, &nd, &ec))
			{
				bc[id * 5 + 0] = (int)ec; bc[id * 5 + 1] = (int)nd;
				bc[id * 5 + 2] = nv[0]; bc[id * 5 + 3] = nv[1]; bc[id * 5 + 4] = nv[2];
			}
		}
		write_raw(Tag + ".wall.i32", bc.data(), bc.size() * sizeof(int));

		FILE *fh = fopen((Tag + ".meta.txt").c_str(), "w");
		double rho_out = L_RHOIN;
#ifdef L_PRESSURE_DELTA
		rho_out += GridUnits::pd2dlbm(L_PRESSURE_DELTA, this);   /* optimised.cpp:343-345 */
#endif
		fprintf(fh, "version=%s\ndims=%d\nQ=%d\nN=%d\nM=%d\nK=%d\n", LUMA_VERSION, L_DIMS, L_NUM_VELS, N_lim, M_lim, K_lim);
		fprintf(fh, "omega=%.17g\nnu=%.17g\ndt=%.17g\ndh=%.17g\ngravity=%.17g\nuref=%.17g\nrho_out=%.17g\ncs=%.17g\n",
			omega, nu, dt, dh, gravity, uref, rho_out, cs);
		fclose(fh);
	}
	write_raw(Tag + ".f.f64", &f[0], f.size() * sizeof(double));
	write_raw(Tag + ".rho.f64", &rho[0], rho.size() * sizeof(double));
	write_raw(Tag + ".u.f64", &u[0], u.size() * sizeof(double));
#ifdef L_COMPUTE_TIME_AVERAGED_QUANTITIES
	/* time-averaged statistics of _LBM_macro_opt (optimised.cpp:895-917) */
	write_raw(Tag + ".rho_timeav.f64", &rho_timeav[0], rho_timeav.size() * sizeof(double));
	write_raw(Tag + ".ui_timeav.f64", &ui_timeav[0], ui_timeav.size() * sizeof(double));
	write_raw(Tag + ".uiuj_timeav.f64", &uiuj_timeav[0], uiuj_timeav.size() * sizeof(double));
#endif
	FILE *fh = fopen((Tag + ".scalars.txt").c_str(), "w");
	ObjectManager *om = ObjectManager::getInstance();
	fprintf(fh, "t=%d\nomega=%.17g\nFx=%.17g\nFy=%.17g\nFz=%.17g\n", t, omega,
		om->*luma_oracle_peek(PeekFx()), om->*luma_oracle_peek(PeekFy()), om->*luma_oracle_peek(PeekFz()));
	fclose(fh);
}

/* ObjectManager::GeomPacked is a private nested type (inc/ObjectManager.h:51); its name cannot be
 * written here, but the type can be deduced from the public member that takes it. */
template <typename Geom>
static void label_body(ObjectManager *om, GridObj *g, PCpts *pts, void (ObjectManager::*add)(GridObj *, Geom *, PCpts *))
{
	Geom geom;
	geom.onGridLev = 0; geom.onGridReg = 0;
	(om->*add)(g, &geom, pts);
}

static GridObj *build_case()
{
	GridManager *gm = GridManager::getInstance();          /* main_lbm.cpp:166 */
	GridObj *Grids = new GridObj(0);                        /* main_lbm.cpp:222 -> LBM_initGrid */
	gm->setGridHierarchy(Grids);
	ObjectManager *om = ObjectManager::getInstance(Grids); /* main_lbm.cpp:273 */
#ifdef LUMA_ORACLE_BOX
	/* Bounce-back body: the reference labels bodies from a point cloud
	 * (io_readInCloud -> addBouncebackObject, src/ObjectManager.cpp:309-345).  The cloud file is
	 * replaced by the cell centres of an index box; the labelling call is the reference's own. */
	{
		const int b[6] = LUMA_ORACLE_BOX;
		PCpts pts;
		for (int i = b[0]; i < b[1]; ++i) for (int j = b[2]; j < b[3]; ++j) for (int k = b[4]; k < b[5]; ++k)
		{
			pts.x.push_back(Grids->XPos[i]); pts.y.push_back(Grids->YPos[j]);
			pts.z.push_back(L_DIMS == 3 ? Grids->ZPos[k] : 0.0);
			pts.id.push_back((int)pts.id.size());
		}
		label_body(om, Grids, &pts, &ObjectManager::addBouncebackObject);
	}
#endif
	(void)om;
	return Grids;
}

int main(int argc, char **argv)
{
	if (argc < 4) { fprintf(stderr, "usage: %s dump <outdir> <steps,csv> | bench <warmup> <steps>\n", argv[0]); return 1; }
	std::string mode = argv[1];
	std::string outdir = (mode == "dump") ? argv[2] : "/tmp";
	if (mode == "dump") mkdir(outdir.c_str(), 0777);
	std::ofstream logfile((outdir + "/luma_ref_log.out").c_str());
	GridUtils::logfile = &logfile;
	GridUtils::path_str = outdir;

	GridObj *Grids = build_case();

	if (mode == "dump")
	{
		std::vector<int> snaps;
		std::stringstream ss(argv[3]); std::string tok;
		while (std::getline(ss, tok, ',')) snaps.push_back(atoi(tok.c_str()));
		Grids->io_lite(-1.0, outdir + "/init");
		size_t next = 0; int last = snaps.empty() ? 0 : snaps.back();
		while (Grids->t < last)
		{
			Grids->LBM_multi_opt();                          /* main_lbm.cpp:441 */
			if (next < snaps.size() && Grids->t == snaps[next])
			{
#ifdef LUMA_DROPIN
				Grids->LBM_multi_opt(-1);                    /* LUMA_B200_SYNC_HOST: refresh the host arrays */
#endif
				Grids->io_lite((double)Grids->t, outdir + "/t" + std::to_string(Grids->t));
				++next;
			}
		}
		return 0;
	}
	else if (mode == "bench")
	{
		int warm = atoi(argv[2]), steps = atoi(argv[3]);
		double cells = (double)Grids->N_lim * Grids->M_lim * Grids->K_lim;
		int threads = 1;
#if defined(_OPENMP) && defined(L_ENABLE_OPENMP)
		threads = omp_get_max_threads();
#endif
#ifdef LUMA_DROPIN
		/* The unmodified host loop stepping on the GPU: LBM_multi_opt() once per step (src/main_lbm.cpp:441).
		 *   first call  = case description + luma_b200_create + luma_b200_upload + step 1 (queued)
		 *   steps       = `steps` calls, then a device sync (the calls themselves never wait)
		 *   download    = rho,u into the GridObj arrays (what the IO points of main_lbm.cpp:449-561 need)
		 * e2e = upload + steps + download (create -- CUDA context start-up and allocations -- is process start-up and
		 * reported separately). */
		auto c0 = std::chrono::steady_clock::now();
		Grids->LBM_multi_opt();
		double first_call = std::chrono::duration<double>(std::chrono::steady_clock::now() - c0).count();
		double create_s = 0.0, upload_s = 0.0;
		luma_b200_t *h = luma_b200_shim_handle(&create_s, &upload_s);
		for (int s = 1; s < warm; ++s) Grids->LBM_multi_opt();
		luma_b200_sync(h);
		auto t0 = std::chrono::steady_clock::now();
		for (int s = 0; s < steps; ++s) Grids->LBM_multi_opt();
		double host_calls = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
		luma_b200_sync(h);
		double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
		auto d0 = std::chrono::steady_clock::now();
		Grids->LBM_multi_opt(-2);                               /* LUMA_B200_SYNC_MACRO: rho,u to the host arrays */
		double down = std::chrono::duration<double>(std::chrono::steady_clock::now() - d0).count();
		LumaStats st;
		luma_b200_stats(h, &st);
		double e2e = upload_s + secs + down;
		printf("{\"mlups\": %.6f, \"seconds\": %.6f, \"steps\": %d, \"warmup\": %d, \"cells\": %.0f, \"threads\": %d, "
			"\"N\": %d, \"M\": %d, \"K\": %d, \"Q\": %d, \"omega\": %.17g, \"first_call_seconds\": %.6f, \"create_seconds\": %.6f, "
			"\"upload_seconds\": %.6f, \"download_seconds\": %.6f, \"host_seconds_in_calls\": %.6f, \"per_call_us\": %.3f, "
			"\"seconds_e2e\": %.6f, \"mlups_e2e\": %.6f, \"graph_launches\": %lld, \"kernel_launches\": %lld}\n",
			cells * steps / secs / 1e6, secs, steps, warm, cells, threads,
			Grids->N_lim, Grids->M_lim, Grids->K_lim, (int)L_NUM_VELS, Grids->omega, first_call, create_s, upload_s, down,
			host_calls, 1e6 * host_calls / steps, e2e, cells * steps / e2e / 1e6, (long long)st.graph_launches, (long long)st.kernel_launches);
		return 0;
#else
		for (int s = 0; s < warm; ++s) Grids->LBM_multi_opt();
		auto t0 = std::chrono::steady_clock::now();
		for (int s = 0; s < steps; ++s) Grids->LBM_multi_opt();
		double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
		printf("{\"mlups\": %.6f, \"seconds\": %.6f, \"steps\": %d, \"warmup\": %d, \"cells\": %.0f, \"threads\": %d, "
			"\"N\": %d, \"M\": %d, \"K\": %d, \"Q\": %d, \"omega\": %.17g}\n",
			cells * steps / secs / 1e6, secs, steps, warm, cells, threads,
			Grids->N_lim, Grids->M_lim, Grids->K_lim, (int)L_NUM_VELS, Grids->omega);
		return 0;
#endif
	}
	fprintf(stderr, "unknown mode %s\n", mode.c_str());
	return 1;
}
