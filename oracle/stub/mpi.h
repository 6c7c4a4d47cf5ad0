/* Single-rank stand-in for <mpi.h>, used ONLY to compile the unmodified reference
 * sources (/root/reference/StePS/src/*.cc) into oracle/_ref/ as a test oracle.
 * Test infrastructure: never linked into the product library.
 * The reference needs the type MPI_Status in every TU (global_variables.h:70) and the
 * point-to-point / Bcast calls of step.cc:183-228 and main.cc; with one rank they are no-ops. */
#ifndef STEPS_B200_ORACLE_STUB_MPI_H
#define STEPS_B200_ORACLE_STUB_MPI_H
typedef struct { int source, tag, error; } MPI_Status;
typedef int MPI_Comm;
typedef int MPI_Datatype;
#define MPI_COMM_WORLD 0
#define MPI_INT 1
#define MPI_FLOAT 2
#define MPI_DOUBLE 3
#define MPI_CHAR 4
#define MPI_SUCCESS 0
static inline int MPI_Init(int *, char ***) { return 0; }
static inline int MPI_Finalize(void) { return 0; }
static inline int MPI_Comm_size(MPI_Comm, int *n) { *n = 1; return 0; }
static inline int MPI_Comm_rank(MPI_Comm, int *r) { *r = 0; return 0; }
static inline int MPI_Bcast(void *, int, MPI_Datatype, int, MPI_Comm) { return 0; }
static inline int MPI_Send(const void *, int, MPI_Datatype, int, int, MPI_Comm) { return 0; }
static inline int MPI_Recv(void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Status *) { return 0; }
static inline int MPI_Barrier(MPI_Comm) { return 0; }
static inline int MPI_Abort(MPI_Comm, int) { return 0; }
#endif
