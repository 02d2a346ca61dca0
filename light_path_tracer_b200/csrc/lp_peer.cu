// lp_peer.cu — completion flags for frames assembled through NVLink peer memory
// (lightpath.h: lp_peer_signal / lp_peer_wait; SURVEY.md 8e — the reference has no
// distribution).
//
// Every rank's render kernel stores its row tile straight into the root GPU's frame; what is
// left of the "gather" is ordering: the root may read frame e only after every rank's kernel
// has finished, and a rank may overwrite a frame buffer only after the root has consumed it.
// Both are one 8-byte flag per (rank, direction): a stream-ordered one-thread kernel stores an
// epoch with release semantics at system scope (a kernel boundary has already made the render
// kernel's peer stores visible), and a one-thread kernel on the other side spins with acquire
// loads.  No collective, no host round trip; ~2 us per call.
#include "lp_internal.cuh"

#define LP_PEER_MAX_FLAGS 16

struct PeerFlagList {
    unsigned long long *p[LP_PEER_MAX_FLAGS];
};

__global__ void lp_peer_signal_kernel(const PeerFlagList flags, int n, unsigned long long value)
{
    const int i = threadIdx.x;
    if (i >= n) return;
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flags.p[i]), "l"(value) : "memory");
}

__global__ void lp_peer_wait_kernel(const unsigned long long *flags, int n, unsigned long long value,
                                    unsigned long long timeout_ns, int *timed_out)
{
    const int i = threadIdx.x;
    if (i >= n) return;
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    unsigned spins = 0;
    while (true) {
        unsigned long long v;
        asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flags + i) : "memory");
        if (v >= value) break;
        if ((++spins & 0x3ffu) == 0u) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (t - t0 > timeout_ns) {
                if (timed_out) atomicExch(timed_out, 1);
                break;
            }
        }
        __nanosleep(64);
    }
    __threadfence_system();
}

extern "C" int lp_peer_signal(uint64_t *const *h_flags, int32_t n_flags, uint64_t value, void *stream)
{
    if (n_flags < 0 || n_flags > LP_PEER_MAX_FLAGS || (n_flags > 0 && !h_flags)) return LP_ERR_INVALID_ARG;
    if (n_flags == 0) return LP_OK;
    PeerFlagList l = {};
    for (int i = 0; i < n_flags; ++i) {
        if (!h_flags[i]) return LP_ERR_INVALID_ARG;
        l.p[i] = (unsigned long long *)h_flags[i];
    }
    lp_peer_signal_kernel<<<1, LP_PEER_MAX_FLAGS, 0, (cudaStream_t)stream>>>(l, n_flags, (unsigned long long)value);
    return lp_check_launch();
}

extern "C" int lp_peer_wait(const uint64_t *flags, int32_t n_flags, uint64_t value,
                            uint32_t timeout_ms, int32_t *timed_out, void *stream)
{
    if (n_flags < 0 || n_flags > 1024 || (n_flags > 0 && !flags)) return LP_ERR_INVALID_ARG;
    if (n_flags == 0) return LP_OK;
    const unsigned long long ns = (timeout_ms ? (unsigned long long)timeout_ms : 10000ull) * 1000000ull;
    lp_peer_wait_kernel<<<1, n_flags, 0, (cudaStream_t)stream>>>((const unsigned long long *)flags, n_flags,
                                                               (unsigned long long)value, ns, timed_out);
    return lp_check_launch();
}
