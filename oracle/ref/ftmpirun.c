/*
 * ftmpirun.c -- launcher for the process-per-rank MPI shim (ftmpi.c).
 * TEST INFRASTRUCTURE ONLY.   usage: ftmpirun -np P program [args...]
 * Creates the shared ring file in /dev/shm, forks P ranks, waits for them and
 * returns the largest exit status (MPI_Abort codes propagate).
 */
#define _GNU_SOURCE
#include <fcntl.h>
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <sys/types.h>
#include <sys/wait.h>
#include <unistd.h>

int main(int argc, char **argv) {
  if (argc < 4 || strcmp(argv[1], "-np") != 0) {
    fprintf(stderr, "usage: %s -np P program [args...]\n", argv[0]);
    return 2;
  }
  int P = atoi(argv[2]);
  if (P < 1 || P > 64) { fprintf(stderr, "ftmpirun: bad -np\n"); return 2; }
  const char *rm = getenv("FTMPI_RING_MB");
  size_t ring = (size_t)(rm ? atoi(rm) : 64) << 20;
  size_t stride = 128 + ring; /* sizeof(ft_ring_ctl) == 128 */
  size_t total = 4096 + stride * (size_t)P * P;
  char path[256];
  snprintf(path, sizeof path, "/dev/shm/ftmpi_%d", (int)getpid());
  int fd = open(path, O_RDWR | O_CREAT | O_TRUNC, 0600);
  if (fd < 0 || ftruncate(fd, (off_t)total) != 0) { perror("ftmpirun: shm"); return 2; }
  close(fd);
  pid_t *pids = (pid_t *)calloc(P, sizeof(pid_t));
  for (int r = 0; r < P; ++r) {
    pid_t pid = fork();
    if (pid == 0) {
      char b[32];
      snprintf(b, sizeof b, "%d", r); setenv("FTMPI_RANK", b, 1);
      snprintf(b, sizeof b, "%d", P); setenv("FTMPI_SIZE", b, 1);
      setenv("FTMPI_SHM", path, 1);
      execvp(argv[3], &argv[3]);
      perror("ftmpirun: exec");
      _exit(127);
    }
    pids[r] = pid;
  }
  int worst = 0, left = P;
  while (left > 0) {
    int st = 0;
    pid_t p = wait(&st);
    if (p < 0) break;
    --left;
    int code = WIFEXITED(st) ? WEXITSTATUS(st) : 128 + WTERMSIG(st);
    if (code > worst) worst = code;
    if (code != 0) /* a rank died: the others will see the abort flag or hang; give them a moment, then kill */
      for (int r = 0; r < P; ++r) if (pids[r] != p) kill(pids[r], SIGTERM);
  }
  unlink(path);
  free(pids);
  return worst;
}
