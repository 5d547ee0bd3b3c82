// tests/cpp/test_xchg.cpp -- the multi-GPU exchange from C++, no Python and no communicator in
// the library: two processes (fork), one engine each (on two GPUs when the box has them, else
// both on GPU 0 -- CUDA IPC works between processes on one device too), the 64-byte handles and
// the barriers travel over a socketpair the way MPI_Allgather / MPI_Barrier would carry them
// (INTEGRATION.md section 4).  Each rank runs its own members with hx_run_exchange and then
// checks that its gather block holds the PEER's trajectories, against numbers the peer fetched
// for itself.   usage: test_xchg <hector_ssp245.ini>
#include <cuda_runtime.h>
#include <sys/socket.h>
#include <sys/wait.h>
#include <unistd.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "hector_b200.h"

static int sock = -1;
static void send_all(const void *p, size_t n) {
  const char *c = (const char *)p;
  while (n) { ssize_t k = write(sock, c, n); if (k <= 0) _exit(3); c += k; n -= (size_t)k; }
}
static void recv_all(void *p, size_t n) {
  char *c = (char *)p;
  while (n) { ssize_t k = read(sock, c, n); if (k <= 0) _exit(3); c += k; n -= (size_t)k; }
}
static void barrier() { char x = 1, y = 0; send_all(&x, 1); recv_all(&y, 1); }
#define CHECK(call)                                                                   \
  do {                                                                                \
    int rc_ = (call);                                                                 \
    if (rc_ != HX_OK) { std::printf("FAILED rank %d: %s -> %d (%s)\n", rank, #call, rc_, hx_last_error(h)); return 1; } \
  } while (0)

static int run_rank(int rank, const char *ini) {
  const int M = 300, NR = 2;
  int ndev = 0;
  cudaGetDeviceCount(&ndev);
  hx_handle h = nullptr;
  const char *inis[1] = {ini};
  if (hx_create_from_ini(inis, 1, M, ndev > 1 ? rank : 0, 0, &h) != HX_OK) {
    std::printf("FAILED rank %d: hx_create_from_ini: %s\n", rank, hx_last_error(nullptr));
    return 1;
  }
  std::vector<double> S(M);
  for (int i = 0; i < M; ++i) S[i] = 2.0 + 0.01 * i + 1.5 * rank;
  CHECK(hx_set_param(h, "S", S.data(), M));
  const char *outs[2] = {"CO2_concentration", "global_tas"};
  CHECK(hx_select_outputs(h, 2, outs));
  CHECK(hx_prepare(h));
  unsigned char mine[64], all[128];
  CHECK(hx_xchg_create(h, NR, rank, mine, nullptr));
  unsigned char theirs[64];
  send_all(mine, 64);
  recv_all(theirs, 64);
  std::memcpy(all + 64 * rank, mine, 64);
  std::memcpy(all + 64 * (1 - rank), theirs, 64);
  CHECK(hx_xchg_open(h, NR, all));
  for (int rep = 0; rep < 2; ++rep) { /* twice: the blocks are reused */
    barrier();
    CHECK(hx_reset(h));
    CHECK(hx_run_exchange(h, -1.0));
    barrier();
  }
  /* what this rank computed for itself, sent to the peer for comparison */
  const double dates[2] = {2000.0, 2300.0};
  std::vector<double> own((size_t)M * 2 * 2), peer_says((size_t)M * 2 * 2);
  CHECK(hx_fetch(h, "CO2_concentration", dates, 2, own.data()));
  CHECK(hx_fetch(h, "global_tas", dates, 2, own.data() + (size_t)M * 2));
  send_all(own.data(), own.size() * sizeof(double));
  recv_all(peer_says.data(), peer_says.size() * sizeof(double));
  const double *block = nullptr, *dev = nullptr;
  int64_t per_rank = 0, stride = 0;
  int32_t ny = 0;
  CHECK(hx_xchg_block(h, &block, &per_rank));
  CHECK(hx_output_device(h, "CO2_concentration", &dev, &stride, &ny));
  std::vector<double> host((size_t)per_rank * NR);
  if (cudaMemcpy(host.data(), block, host.size() * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess) {
    std::printf("FAILED rank %d: cudaMemcpy of the gather block\n", rank);
    return 1;
  }
  int bad = 0;
  for (int who = 0; who < NR; ++who) {
    const std::vector<double> &ref = who == rank ? own : peer_says;
    for (int v = 0; v < 2; ++v)
      for (int k = 0; k < 2; ++k) {
        const int row = (int)dates[k] - 1746;
        for (int i = 0; i < M; ++i) {
          const double got = host[(size_t)who * per_rank + ((size_t)v * ny + row) * stride + i];
          const double want = ref[(size_t)v * M * 2 + (size_t)i * 2 + k];
          if (!(got == want)) ++bad;
        }
      }
  }
  barrier(); /* nobody tears its block down while the peer still reads */
  hx_destroy(h);
  std::printf("rank %d: gather block %s (%d mismatches)\n", rank, bad ? "WRONG" : "matches both ranks' own fetches", bad);
  return bad ? 1 : 0;
}

int main(int argc, char **argv) {
  if (argc < 2) return 2;
  int sv[2];
  if (socketpair(AF_UNIX, SOCK_STREAM, 0, sv) != 0) return 2;
  const pid_t child = fork(); /* before any CUDA call */
  if (child == 0) {
    close(sv[0]);
    sock = sv[1];
    _exit(run_rank(1, argv[1]));
  }
  close(sv[1]);
  sock = sv[0];
  const int rc0 = run_rank(0, argv[1]);
  int st = 0;
  waitpid(child, &st, 0);
  const int rc1 = WIFEXITED(st) ? WEXITSTATUS(st) : 9;
  std::printf("%s\n", (rc0 == 0 && rc1 == 0) ? "XCHG_OK" : "XCHG_FAILED");
  return (rc0 == 0 && rc1 == 0) ? 0 : 1;
}
