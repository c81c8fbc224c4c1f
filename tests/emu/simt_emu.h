/* TEST INFRASTRUCTURE - NOT PRODUCT CODE.
 *
 * Minimal SIMT emulation so that the product's device header (mve_b200/csrc/patch_opt.cuh) can be compiled by g++ and
 * executed on the CPU: one warp = 32 host threads that each run the scalar device code; the warp collectives
 * (__shfl*_sync, __ballot_sync, __any/__all_sync) exchange values through a per-warp scratch area guarded by a
 * barrier.  The device code only ever calls collectives convergently with the full mask (or, for sub-warp variants,
 * with the mask of an aligned lane group), which is what this emulation supports.
 *
 * Differences from the GPU: rcp/rsqrt are exact divisions (the header selects them under B200MVS_HOST_EMU), libm
 * instead of the CUDA math library, no FMA contraction guarantees.  It checks the LOGIC of the kernel (state machine,
 * reductions, lane-distributed arrays, view selection), not its last-bit numerics.
 */
#ifndef SIMT_EMU_H
#define SIMT_EMU_H

#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __launch_bounds__(...)

struct uchar4 { unsigned char x, y, z, w; };
struct float4 { float x, y, z, w; };
struct int4 { int x, y, z, w; };
struct uint4 { unsigned x, y, z, w; };
inline int min(int a, int b) { return a < b ? a : b; }
inline float4 make_float4(float x, float y, float z, float w) { float4 r = {x, y, z, w}; return r; }
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { uint4 r = {x, y, z, w}; return r; }
inline uchar4 make_uchar4(unsigned char x, unsigned char y, unsigned char z, unsigned char w) { uchar4 r = {x, y, z, w}; return r; }

template <typename T> inline T __ldg(const T* p) { return *p; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
inline unsigned __float_as_uint(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }

namespace simt_emu {

struct Barrier {                       /* sense-reversing barrier for a fixed number of lanes */
    std::atomic<int> count{0};
    std::atomic<int> sense{0};
    int n = 32;
    void wait(int& local_sense)
    {
        local_sense ^= 1;
        if (count.fetch_add(1, std::memory_order_acq_rel) == n - 1) {
            count.store(0, std::memory_order_relaxed);
            sense.store(local_sense, std::memory_order_release);
        } else {
            int spins = 0;
            while (sense.load(std::memory_order_acquire) != local_sense)
                if (++spins > 64) { std::this_thread::yield(); spins = 0; }
        }
    }
};

struct Group {                         /* lanes that execute collectives together (a warp, or an aligned sub-warp group) */
    Barrier bar;
    uint64_t slot[32];
};

struct Warp {
    Group full;                        /* mask 0xffffffff */
    Group sub8[4];                     /* masks 0xff << 8g */
};

extern thread_local Warp* t_warp;
extern thread_local int t_lane;
extern thread_local int t_sense_full;
extern thread_local int t_sense_sub;

inline Group& group_of(unsigned mask, int*& sense)
{
    if (mask == 0xffffffffu) { sense = &t_sense_full; return t_warp->full; }
    sense = &t_sense_sub;
    return t_warp->sub8[t_lane >> 3];
}

/* every participating lane publishes `v`, then reads whatever it needs through `pick(slots)` */
template <typename T, typename F> inline T exchange(unsigned mask, T v, F pick)
{
    static_assert(sizeof(T) <= 8, "shuffle payload");
    int* sense;
    Group& g = group_of(mask, sense);
    uint64_t raw = 0;
    std::memcpy(&raw, &v, sizeof(T));
    g.slot[t_lane] = raw;
    g.bar.wait(*sense);
    const int src = pick() & 31;
    uint64_t got = g.slot[src];
    T out;
    std::memcpy(&out, &got, sizeof(T));
    g.bar.wait(*sense);                /* nobody overwrites a slot before everybody has read */
    return out;
}

} // namespace simt_emu

template <typename T> inline T __shfl_sync(unsigned mask, T v, int src, int width = 32)
{
    const int lane = simt_emu::t_lane;
    return simt_emu::exchange(mask, v, [=]() { return (lane & ~(width - 1)) | (src & (width - 1)); });
}
template <typename T> inline T __shfl_xor_sync(unsigned mask, T v, int lanemask, int width = 32)
{
    (void)width;
    const int lane = simt_emu::t_lane;
    return simt_emu::exchange(mask, v, [=]() { return lane ^ lanemask; });
}
template <typename T> inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32)
{
    const int lane = simt_emu::t_lane;
    return simt_emu::exchange(mask, v, [=]() { return ((lane & (width - 1)) >= (int)delta) ? lane - (int)delta : lane; });
}
inline unsigned __ballot_sync(unsigned mask, int pred)
{
    using namespace simt_emu;
    int* sense;
    Group& g = group_of(mask, sense);
    g.slot[t_lane] = pred ? 1u : 0u;
    g.bar.wait(*sense);
    unsigned r = 0;
    for (int l = 0; l < 32; ++l) if (((mask >> l) & 1u) && g.slot[l]) r |= 1u << l;
    g.bar.wait(*sense);
    return r;
}
inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) == mask; }

#endif
