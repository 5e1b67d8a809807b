// Peer-memory exchange of the scale_noise statistics between the GPUs of one node (NVLink 5 /
// NVSwitch), replacing a latency-bound NCCL all-reduce of two doubles per sampler step.
//
// Reference behaviour being preserved: scale_noise (py/utils.py:100-106) reduces over the WHOLE
// batch, so a batch-sharded run must combine every rank's (sum, sum^2) before normalising
// (caveat 1). Each rank owns a "mailbox" in its own HBM, mapped into every
// peer with CUDA IPC. After its moments pass a rank STORES its two partial sums straight into the
// mailbox of every peer over NVLink (payload, system fence, epoch flag); the consuming step kernel
// spins on the N flags of its LOCAL mailbox and adds the N partials in rank order (deterministic).
// No host round trip, no collective launch: ~2 us instead of ~45 us per exchange.
//
// Mailbox layout: double box[2 (epoch parity)][SONAR_PEER_MAX_RANKS][4] = {sum, sum^2, epoch, pad},
// followed by a table region double tab[2][SONAR_PEER_MAX_RANKS][2 + SONAR_PEER_TABLE_MAX] =
// {epoch, n, payload...} used by sonar_peer_allreduce_table (the (K, 2) look-ahead statistics of a
// whole sampler run: one exchange per run instead of one per step).
#include "common.cuh"
#include "../../include/sonar_b200.h"

namespace sonar {

struct PeerTargets {
  double* box[SONAR_PEER_MAX_RANKS];  // mailbox base of every rank, as mapped in THIS process
};

__global__ void peer_publish_kernel(PeerTargets targets, const double* __restrict__ local_sums, int rank, int world,
                                    double epoch) {
  const int dst = threadIdx.x;
  if (dst >= world) return;
  const int parity = ((long long)epoch) & 1;
  volatile double* slot = targets.box[dst] + ((size_t)parity * SONAR_PEER_MAX_RANKS + rank) * 4;
  slot[0] = local_sums[0];
  slot[1] = local_sums[1];
  __threadfence_system();  // payload visible system-wide before the flag
  slot[2] = epoch;
}

constexpr size_t kPairDoubles = 2 * SONAR_PEER_MAX_RANKS * 4;
constexpr size_t kTableSlot = 2 + SONAR_PEER_TABLE_MAX;
constexpr size_t kMailboxDoubles = kPairDoubles + 2 * SONAR_PEER_MAX_RANKS * kTableSlot;

// One CTA: (1) store this rank's table into every rank's mailbox over NVLink, fence, raise the epoch
// flags; (2) wait until every rank's table of this epoch has landed in the LOCAL mailbox; (3) replace
// the local table by the sum over ranks, added in rank order so that all ranks hold identical bits.
__global__ void __launch_bounds__(512)
peer_allreduce_table_kernel(PeerTargets targets, double* __restrict__ table, int n, int rank, int world, double epoch) {
  const int parity = ((long long)epoch) & 1;
  const size_t slot_of_me = kPairDoubles + ((size_t)parity * SONAR_PEER_MAX_RANKS + rank) * kTableSlot;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double v = table[i];
    for (int dst = 0; dst < world; ++dst) ((volatile double*)targets.box[dst])[slot_of_me + 2 + i] = v;
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x < world) {
    volatile double* slot = targets.box[threadIdx.x] + slot_of_me;
    slot[1] = (double)n;
    __threadfence_system();
    slot[0] = epoch;
  }
  // consume: thread r waits for rank r's flag in the local mailbox
  volatile double* local = targets.box[rank];
  if (threadIdx.x < world) {
    const size_t s = kPairDoubles + ((size_t)parity * SONAR_PEER_MAX_RANKS + threadIdx.x) * kTableSlot;
    wait_for_epoch(&local[s], epoch);
    __threadfence_system();
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double acc = 0.0;
    for (int r = 0; r < world; ++r)
      acc += local[kPairDoubles + ((size_t)parity * SONAR_PEER_MAX_RANKS + r) * kTableSlot + 2 + i];
    table[i] = acc;
  }
}

}  // namespace sonar

extern "C" {

int sonar_peer_alloc(void** mailbox_out) {
  if (mailbox_out == nullptr) return (int)cudaErrorInvalidValue;
  const size_t bytes = sonar::kMailboxDoubles * sizeof(double);
  SONAR_CUDA_TRY(cudaMalloc(mailbox_out, bytes));
  SONAR_CUDA_TRY(cudaMemset(*mailbox_out, 0, bytes));
  return 0;
}

int sonar_peer_free(void* mailbox) { return (int)cudaFree(mailbox); }

int sonar_peer_get_handle(const void* mailbox, unsigned char* handle64_host) {
  cudaIpcMemHandle_t h;
  SONAR_CUDA_TRY(cudaIpcGetMemHandle(&h, const_cast<void*>(mailbox)));
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  for (int i = 0; i < 64; ++i) handle64_host[i] = reinterpret_cast<unsigned char*>(&h)[i];
  return 0;
}

int sonar_peer_open_handle(const unsigned char* handle64_host, void** mapped_out) {
  cudaIpcMemHandle_t h;
  for (int i = 0; i < 64; ++i) reinterpret_cast<unsigned char*>(&h)[i] = handle64_host[i];
  SONAR_CUDA_TRY(cudaIpcOpenMemHandle(mapped_out, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}

int sonar_peer_close_handle(void* mapped) { return (int)cudaIpcCloseMemHandle(mapped); }

int sonar_peer_publish_sums(void* const* mailboxes_host, int rank, int world, const double* local_sums, double epoch,
                            void* stream) {
  if (world < 1 || world > SONAR_PEER_MAX_RANKS || rank < 0 || rank >= world) return (int)cudaErrorInvalidValue;
  sonar::PeerTargets t;
  for (int r = 0; r < SONAR_PEER_MAX_RANKS; ++r) t.box[r] = r < world ? (double*)mailboxes_host[r] : nullptr;
  sonar::peer_publish_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(t, local_sums, rank, world, epoch);
  SONAR_LAUNCH_CHECK();
  return 0;
}

int sonar_peer_allreduce_table(void* const* mailboxes_host, int rank, int world, double* table, int n, double epoch,
                               void* stream) {
  if (world < 1 || world > SONAR_PEER_MAX_RANKS || rank < 0 || rank >= world || table == nullptr || n < 0 ||
      n > SONAR_PEER_TABLE_MAX)
    return (int)cudaErrorInvalidValue;
  if (n == 0) return 0;
  sonar::PeerTargets t;
  for (int r = 0; r < SONAR_PEER_MAX_RANKS; ++r) t.box[r] = r < world ? (double*)mailboxes_host[r] : nullptr;
  sonar::peer_allreduce_table_kernel<<<1, 512, 0, (cudaStream_t)stream>>>(t, table, n, rank, world, epoch);
  SONAR_LAUNCH_CHECK();
  return 0;
}

}  // extern "C"
