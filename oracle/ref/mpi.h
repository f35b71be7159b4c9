/*
 * mpi.h -- minimal MPI interface for building the FemTech reference (and its
 * vendored ParMETIS) inside this repository's oracle harness.  TEST
 * INFRASTRUCTURE ONLY.  Neither this container nor the GPU box has an MPI
 * installation; this header plus ftmpi.c provide the ~45 entry points the
 * reference uses (list: SURVEY.md section 8c).  Handles are plain ints.
 */
#ifndef FTMPI_MPI_H
#define FTMPI_MPI_H
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h> /* src/io/parallel_log.cpp relies on this transitively */
#ifdef __cplusplus
extern "C" {
#endif
typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Op;
typedef int MPI_Request;
typedef int MPI_Info;
typedef int MPI_Fint;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR, ftmpi_bytes; } MPI_Status;
typedef struct ftmpi_file *MPI_File;

#define MPI_COMM_WORLD 0
#define MPI_COMM_NULL (-1)
#define MPI_SUCCESS 0
#define MPI_UNDEFINED (-32766)
#define MPI_ANY_TAG (-1)
#define MPI_REQUEST_NULL (-1)

#define MPI_DATATYPE_NULL 0
#define MPI_CHAR 1
#define MPI_BYTE 2
#define MPI_SHORT 3
#define MPI_INT 4
#define MPI_FLOAT 5
#define MPI_LONG 6
#define MPI_UNSIGNED 7
#define MPI_DOUBLE 8
#define MPI_LONG_LONG_INT 9
#define MPI_LONG_LONG 9
#define MPI_UNSIGNED_LONG 10
#define MPI_DOUBLE_INT 12
#define MPI_2INT 13
#define MPI_FLOAT_INT 14

#define MPI_SUM 1
#define MPI_MIN 2
#define MPI_MAX 3
#define MPI_MINLOC 4
#define MPI_MAXLOC 5
#define MPI_LOR 6
#define MPI_LAND 7
#define MPI_BOR 8

#define MPI_IN_PLACE ((void *)1)
#define MPI_STATUS_IGNORE ((MPI_Status *)0)
#define MPI_STATUSES_IGNORE ((MPI_Status *)0)
#define MPI_INFO_NULL 0
#define MPI_MODE_EXCL 1
#define MPI_MODE_WRONLY 2
#define MPI_MODE_CREATE 4
#define MPI_MODE_APPEND 8

int MPI_Init(int *, char ***);
int MPI_Initialized(int *);
int MPI_Finalize(void);
int MPI_Abort(MPI_Comm, int);
int MPI_Comm_size(MPI_Comm, int *);
int MPI_Comm_rank(MPI_Comm, int *);
int MPI_Comm_dup(MPI_Comm, MPI_Comm *);
int MPI_Comm_free(MPI_Comm *);
int MPI_Comm_split(MPI_Comm, int, int, MPI_Comm *);
int MPI_Barrier(MPI_Comm);
double MPI_Wtime(void);
int MPI_Allreduce(const void *, void *, int, MPI_Datatype, MPI_Op, MPI_Comm);
int MPI_Reduce(const void *, void *, int, MPI_Datatype, MPI_Op, int, MPI_Comm);
int MPI_Scan(const void *, void *, int, MPI_Datatype, MPI_Op, MPI_Comm);
int MPI_Bcast(void *, int, MPI_Datatype, int, MPI_Comm);
int MPI_Alltoall(const void *, int, MPI_Datatype, void *, int, MPI_Datatype, MPI_Comm);
int MPI_Alltoallv(const void *, const int *, const int *, MPI_Datatype, void *,
                  const int *, const int *, MPI_Datatype, MPI_Comm);
int MPI_Allgather(const void *, int, MPI_Datatype, void *, int, MPI_Datatype, MPI_Comm);
int MPI_Allgatherv(const void *, int, MPI_Datatype, void *, const int *,
                   const int *, MPI_Datatype, MPI_Comm);
int MPI_Gather(const void *, int, MPI_Datatype, void *, int, MPI_Datatype, int, MPI_Comm);
int MPI_Gatherv(const void *, int, MPI_Datatype, void *, const int *,
                const int *, MPI_Datatype, int, MPI_Comm);
int MPI_Scatterv(const void *, const int *, const int *, MPI_Datatype, void *,
                 int, MPI_Datatype, int, MPI_Comm);
int MPI_Isend(const void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Request *);
int MPI_Irecv(void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Request *);
int MPI_Send(const void *, int, MPI_Datatype, int, int, MPI_Comm);
int MPI_Recv(void *, int, MPI_Datatype, int, int, MPI_Comm, MPI_Status *);
int MPI_Wait(MPI_Request *, MPI_Status *);
int MPI_Waitall(int, MPI_Request *, MPI_Status *);
int MPI_Get_count(const MPI_Status *, MPI_Datatype, int *);
int MPI_Info_create(MPI_Info *);
int MPI_Info_set(MPI_Info, const char *, const char *);
int MPI_Info_free(MPI_Info *);
int MPI_File_open(MPI_Comm, const char *, int, MPI_Info, MPI_File *);
int MPI_File_close(MPI_File *);
int MPI_File_delete(const char *, MPI_Info);
int MPI_File_write_shared(MPI_File, const void *, int, MPI_Datatype, MPI_Status *);
int MPI_File_write_ordered(MPI_File, const void *, int, MPI_Datatype, MPI_Status *);
#ifdef __cplusplus
}
#endif
#endif
