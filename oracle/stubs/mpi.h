/* Stub <mpi.h> for building the reference's hot-path files without MPI.
 * TEST INFRASTRUCTURE ONLY (oracle/): the RCB force tree never communicates; it only
 * reaches <mpi.h> through Partition.h -> Definition.h -> rru_mpi.h for one
 * Partition::getMyProc() in a printf (reference src/halo_finder/RCBForceTree.cxx:84,503). */
#ifndef HACCSR_ORACLE_STUB_MPI_H
#define HACCSR_ORACLE_STUB_MPI_H
typedef int MPI_Comm;
#define MPI_COMM_WORLD 0
#endif
