/* TEST INFRASTRUCTURE ONLY (oracle build) -- never part of the product path.
 *
 * Serial stand-in for <mpi.h>, so that the LUMA reference sources compile in an
 * image that has no MPI.  The oracle is always built WITHOUT L_BUILD_FOR_MPI,
 * so none of these symbols is reached on the level-0 time-step path
 * (GridUtils::isOnRecvLayer is #ifdef-guarded, src/GridUtils.cpp:1109, and
 * GridObj::LBM_multi_opt only calls MpiManager::mpi_communicate under
 * L_BUILD_FOR_MPI, src/GridObj_ops_lbm_optimised.cpp:186-191).  Every entry
 * point is a no-op returning MPI_SUCCESS; they exist only to satisfy the
 * compiler for translation units such as src/MpiManager.cpp.
 */
/* Syntax check of the MPI branch of luma_b200/host/GridObj_ops_lbm_b200.cpp (tests/test_host_mirror.py): in a real LUMA
 * build L_BUILD_FOR_MPI comes from inc/definitions.h, which inc/stdafx.h includes at :203 -- AFTER errorfcn (:135-149), so
 * that function's MPI lines are never compiled.  The oracle's case header is force-included ahead of everything (and pulls
 * this file in once, oracle/cases/luma_case_tail.h), so the macro is raised on the SECOND inclusion instead: that is
 * inc/stdafx.h:205, right behind definitions.h / GridManager.h and in front of inc/MpiManager.h. */
#if defined(LUMA_SHIM_MPI_SYNTAX) && defined(LUMA_B200_ORACLE_MPI_SHIM_H) && !defined(L_BUILD_FOR_MPI)
#define L_BUILD_FOR_MPI
#endif
#ifndef LUMA_B200_ORACLE_MPI_SHIM_H
#define LUMA_B200_ORACLE_MPI_SHIM_H

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Request;
typedef int MPI_Op;
typedef int MPI_Info;
typedef int MPI_Group;
typedef struct MPI_Status { int MPI_SOURCE; int MPI_TAG; int MPI_ERROR; } MPI_Status;

enum {
	MPI_SUCCESS = 0,
	MPI_COMM_WORLD = 0,
	MPI_COMM_NULL = -1,
	MPI_INFO_NULL = 0,
	MPI_REQUEST_NULL = 0,
	MPI_UNDEFINED = -32766
};
enum { MPI_DOUBLE = 1, MPI_INT, MPI_LONG, MPI_CHAR, MPI_UNSIGNED, MPI_C_BOOL };
enum { MPI_SUM = 1, MPI_MAX, MPI_MIN };
enum { MPI_COMM_TYPE_SHARED = 1 };

#define MPI_STATUS_IGNORE   ((MPI_Status *)0)
#define MPI_STATUSES_IGNORE ((MPI_Status *)0)
#define MPI_IN_PLACE        ((void *)1)

#ifdef __cplusplus
struct LumaOracleMpiNoop {
	template <typename... Args> int operator()(Args...) const { return MPI_SUCCESS; }
};
static const LumaOracleMpiNoop luma_oracle_mpi_noop = LumaOracleMpiNoop();
#define LUMA_ORACLE_MPI_FN(name) static const LumaOracleMpiNoop &name = luma_oracle_mpi_noop;
LUMA_ORACLE_MPI_FN(MPI_Init)          LUMA_ORACLE_MPI_FN(MPI_Finalize)
LUMA_ORACLE_MPI_FN(MPI_Barrier)       LUMA_ORACLE_MPI_FN(MPI_Comm_size)
LUMA_ORACLE_MPI_FN(MPI_Comm_rank)     LUMA_ORACLE_MPI_FN(MPI_Cart_create)
LUMA_ORACLE_MPI_FN(MPI_Cart_coords)   LUMA_ORACLE_MPI_FN(MPI_Cart_rank)
LUMA_ORACLE_MPI_FN(MPI_Bcast)         LUMA_ORACLE_MPI_FN(MPI_Isend)
LUMA_ORACLE_MPI_FN(MPI_Irecv)         LUMA_ORACLE_MPI_FN(MPI_Send)
LUMA_ORACLE_MPI_FN(MPI_Recv)          LUMA_ORACLE_MPI_FN(MPI_Bsend)
LUMA_ORACLE_MPI_FN(MPI_Sendrecv_replace)
LUMA_ORACLE_MPI_FN(MPI_Wait)          LUMA_ORACLE_MPI_FN(MPI_Waitall)
LUMA_ORACLE_MPI_FN(MPI_Comm_split)    LUMA_ORACLE_MPI_FN(MPI_Comm_free)
LUMA_ORACLE_MPI_FN(MPI_Comm_split_type)
LUMA_ORACLE_MPI_FN(MPI_Gather)        LUMA_ORACLE_MPI_FN(MPI_Gatherv)
LUMA_ORACLE_MPI_FN(MPI_Scatter)       LUMA_ORACLE_MPI_FN(MPI_Scatterv)
LUMA_ORACLE_MPI_FN(MPI_Alltoall)      LUMA_ORACLE_MPI_FN(MPI_Alltoallv)
LUMA_ORACLE_MPI_FN(MPI_Allreduce)     LUMA_ORACLE_MPI_FN(MPI_Reduce)
LUMA_ORACLE_MPI_FN(MPI_Allgather)     LUMA_ORACLE_MPI_FN(MPI_Allgatherv)
LUMA_ORACLE_MPI_FN(MPI_Buffer_attach) LUMA_ORACLE_MPI_FN(MPI_Buffer_detach)
LUMA_ORACLE_MPI_FN(MPI_Abort)         LUMA_ORACLE_MPI_FN(MPI_Wtime)
#undef LUMA_ORACLE_MPI_FN
#endif

#endif
