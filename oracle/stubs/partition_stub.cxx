// Definitions of Partition's static members so the reference's RCBForceTree.o links
// without Partition.cxx (which needs a real MPI).  TEST INFRASTRUCTURE ONLY (oracle/).
// Declarations being satisfied: reference src/halo_finder/Partition.h (static members).
#include "Partition.h"
int Partition::numProc = 1;
int Partition::myProc = 0;
int Partition::decompSize[DIMENSION];
int Partition::myPosition[DIMENSION];
int Partition::neighbor[NUM_OF_NEIGHBORS];
int Partition::initialized = 0;
MPI_Comm Partition::cartComm = 0;
