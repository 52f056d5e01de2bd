/* Declarations-only stand-in for <mpi.h>, written for tests/test_refcompat_link.py: MPI is not installed in this
 * image, but compiling the reference's main.c / patterns_over_ranks.c / database_over_ranks.c to OBJECTS only needs
 * the handful of MPI names they use.  Nothing here is ever linked or executed. */
#ifndef APM_TEST_STUB_MPI_H
#define APM_TEST_STUB_MPI_H

/* The reference names Open MPI's predefined objects in an OpenMP shared() clause (database_over_ranks.c:295), so
 * the handles have to be spelled the Open MPI way: pointers to global objects of the runtime. */
struct ompi_communicator_t;
struct ompi_datatype_t;
typedef struct ompi_communicator_t *MPI_Comm;
typedef struct ompi_datatype_t *MPI_Datatype;
extern struct ompi_communicator_t ompi_mpi_comm_world;
extern struct ompi_datatype_t ompi_mpi_int, ompi_mpi_byte;
typedef struct MPI_Status {
    int MPI_SOURCE;
    int MPI_TAG;
    int MPI_ERROR;
} MPI_Status;

#define MPI_COMM_WORLD (&ompi_mpi_comm_world)
#define MPI_BYTE (&ompi_mpi_byte)
#define MPI_INT (&ompi_mpi_int)
#define MPI_ANY_SOURCE (-2)
#define MPI_ANY_TAG (-1)
#define MPI_SUCCESS 0

int MPI_Init(int *argc, char ***argv);
int MPI_Finalize(void);
int MPI_Comm_rank(MPI_Comm comm, int *rank);
int MPI_Comm_size(MPI_Comm comm, int *size);
int MPI_Bcast(void *buffer, int count, MPI_Datatype datatype, int root, MPI_Comm comm);
int MPI_Send(const void *buf, int count, MPI_Datatype datatype, int dest, int tag, MPI_Comm comm);
int MPI_Recv(void *buf, int count, MPI_Datatype datatype, int source, int tag, MPI_Comm comm, MPI_Status *status);
double MPI_Wtime(void);

#endif
