// Device-side signalling and halo pulls of the z-slab driver (bmq3d_mg_*, solver3d.cu): the collectives the
// decomposition needs, done by the library itself over peer-mapped memory (NVLink / NVSwitch) instead of by
// the host's communicator.
//
// Every rank owns one MgSignal block in its own device memory; every other rank maps it (CUDA IPC, or the raw
// pointer inside one process).  Words are only ever written by their owner and polled by the others:
//   arrive      number of the last halo exchange this rank's copy stream has reached -- at that point the
//               kernels that produce the exchanged fields are complete (the stream waited for their event) and
//               the rank's pulls of all earlier exchanges are complete (same stream);
//   red_seq[p]  number of the max-reduction whose contribution sits in red[p] (p = number & 1).
// One exchange is ONE kernel (k_mg_pull): publish `arrive`, wait for the ranks within halo reach, then copy
// every halo segment out of its owner's memory with 128-bit loads over NVLink.  A neighbour-only barrier with
// the properties the all-rank stream barrier of round 1 had: data are complete before
// they are read, and an owner rewrites a buffer only after a later exchange, which its readers reach after
// their reads.
#pragma once
#include <cuda_runtime.h>

namespace bmq {

enum { MG_MAX_WORLD = 16, MG_RED_MAX = 4, MG_MAX_SEG = 72 };

struct MgSignal {
    unsigned arrive;
    unsigned timed_out;          // set by a waiter that gave up (a peer died): results are void
    unsigned red_seq[2];
    float red[2][MG_RED_MAX];
    unsigned pad[20];
};

struct MgWaitList {
    const MgSignal *sig[MG_MAX_WORLD];
    int n;
};

struct MgSeg {
    const void *src;             // in the owner's memory
    void *dst;                   // in this rank's halo
    unsigned long long bytes;    // multiple of 4
};
struct MgSegList {
    MgSeg seg[MG_MAX_SEG];
    int n;
};

// epoch == 0: no barrier (continuation launch of an exchange with more than MG_MAX_SEG segments)
cudaError_t launch_mg_pull(cudaStream_t s, MgSignal *mine, const MgWaitList &wait, unsigned epoch, const MgSegList &segs);
// max over all ranks of n <= MG_RED_MAX non-negative floats, this rank's contribution given by value (`vals`, host) or
// as the result of an earlier kernel in `s` (`dev_vals`, device; wins when not null); host_out (pinned) = the n
// maxima, then the timed_out flag
cudaError_t launch_mg_allreduce_max(cudaStream_t s, MgSignal *mine, const MgWaitList &others, const float *vals, const float *dev_vals,
                                    int n, unsigned seq, float *host_out);

}  // namespace bmq
