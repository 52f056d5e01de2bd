/* Definitions for tests/stub_mpi/mpi.h so that the link-level test can close the link with --no-undefined.
 * Never executed: every entry point aborts. */
#include "mpi.h"

#include <stdlib.h>

struct ompi_communicator_t { int unused; } ompi_mpi_comm_world;
struct ompi_datatype_t { int unused; } ompi_mpi_int, ompi_mpi_byte;

int MPI_Init(int *argc, char ***argv) { (void)argc; (void)argv; abort(); }
int MPI_Finalize(void) { abort(); }
int MPI_Comm_rank(MPI_Comm comm, int *rank) { (void)comm; (void)rank; abort(); }
int MPI_Comm_size(MPI_Comm comm, int *size) { (void)comm; (void)size; abort(); }
int MPI_Bcast(void *buffer, int count, MPI_Datatype datatype, int root, MPI_Comm comm) {
    (void)buffer; (void)count; (void)datatype; (void)root; (void)comm; abort();
}
int MPI_Send(const void *buf, int count, MPI_Datatype datatype, int dest, int tag, MPI_Comm comm) {
    (void)buf; (void)count; (void)datatype; (void)dest; (void)tag; (void)comm; abort();
}
int MPI_Recv(void *buf, int count, MPI_Datatype datatype, int source, int tag, MPI_Comm comm, MPI_Status *status) {
    (void)buf; (void)count; (void)datatype; (void)source; (void)tag; (void)comm; (void)status; abort();
}
double MPI_Wtime(void) { abort(); }
