// Host-side scheduler of the track kernels (fk_tracks_kernel.cuh).
//
// fk walks a tree: joint i needs the global transform of parents[i] (ops/skeleton.py:51-58 does it in index
// order, one joint after the other).  Joints on different branches do not depend on each other, and the
// lane = (frame, row) kernels are bound by the latency of ONE dependent chain per warp.  The schedule below
// turns the tree into T steps of U independent items ("tracks"): the U joints of a step have all their parents
// in earlier steps, so a thread can interleave U chains (instruction-level parallelism instead of more warps,
// which shared memory does not allow).
//
// Unit-time tasks with tree precedence on U identical machines: highest-level-first list scheduling (Hu 1961)
// is optimal, so T is the minimum number of steps for U tracks.  Among the joints picked for a step, one whose
// parent was processed by track u in the previous step is given to track u again: its parent's row is still
// in that track's registers ("carry") and is not re-read from the shared-memory stage.
//
// Item code, one 32-bit word per (step, track):  bits 0-9 joint | bits 10-19 parent | bit 20 carry | bit 21 no-op.
#pragma once
#include <stdint.h>

#include <algorithm>
#include <vector>

#include "../../include/pymotion_b200.h"

namespace pmb {

constexpr int kTrackCap = 1024;  // items (steps x tracks) per program: travels as a 4 KB kernel parameter
constexpr uint32_t kTrackCarry = 1u << 20, kTrackNoop = 1u << 21;

constexpr int kTrackMaxChunks = PMB_MAX_JOINTS / 8;  // windows of 8 joints (one TMA box of quaternions)
struct TrackProgram {
    uint32_t code[kTrackCap];
    uint16_t chunk_first[kTrackMaxChunks + 2];  // first step of every window, then the total step count
};

__host__ __device__ inline uint32_t track_joint(uint32_t c) { return c & 0x3FFu; }
__host__ __device__ inline uint32_t track_parent(uint32_t c) { return (c >> 10) & 0x3FFu; }

// `window` > 0: joints are scheduled window by window in index order (a step only holds joints of one window of
// `window` consecutive joints -- what is resident of the input at any time); chunk_first[w] receives the first step
// of window w and chunk_first[n_windows] the total.  window = 0: one window, the whole skeleton.
// Returns the number of steps T (items are code[t * n_tracks + u]), 0 if the schedule does not fit kTrackCap,
// -1 for a bad parents table (bad_joint set).
inline int build_track_schedule(const int64_t *parents, int n_joints, int n_tracks, uint32_t *code, int *bad_joint = nullptr,
                                int window = 0, uint16_t *chunk_first = nullptr) {
    if (n_joints < 1 || n_joints > PMB_MAX_JOINTS || n_tracks < 1) return -1;
    for (int i = 1; i < n_joints; ++i)
        if (parents[i] < 0 || parents[i] >= i) {
            if (bad_joint) *bad_joint = i;
            return -1;
        }
    const int U = n_tracks;
    // height = joints on the longest path from the joint down to a leaf (the "level" of Hu's algorithm)
    std::vector<int> height(n_joints, 1), step_of(n_joints, -1);
    for (int i = n_joints - 1; i >= 1; --i) {
        const int p = static_cast<int>(parents[i]);
        height[p] = std::max(height[p], height[i] + 1);
    }
    std::vector<int> last(U, -1), pick, ready;  // last[u]: joint track u processed in the previous step
    int done = 0, t = 0;
    const int W = window > 0 ? window : n_joints;
    if (chunk_first) chunk_first[0] = 0;
    // step 0: the root, on track 0, carried from the registers the kernel initialises (identity at global_pos)
    {
        if (U > kTrackCap) return 0;
        for (int u = 0; u < U; ++u) code[u] = kTrackNoop;
        code[0] = 0u | kTrackCarry;
        step_of[0] = 0, last[0] = 0, done = 1, t = 1;
    }
    int w_lo = 0, w_done = 1;  // current window [w_lo, w_lo + W), joints of it already scheduled
    while (done < n_joints) {
        const int w_hi = std::min(n_joints, w_lo + W);
        if (w_done == w_hi - w_lo) {  // window finished: the next one starts at this step
            w_lo = w_hi, w_done = 0;
            if (chunk_first) chunk_first[w_lo / W] = static_cast<uint16_t>(t);
            continue;
        }
        if ((t + 1) * U > kTrackCap) return 0;
        ready.clear();
        for (int i = std::max(1, w_lo); i < w_hi; ++i) {
            const int p = static_cast<int>(parents[i]);
            if (step_of[i] < 0 && step_of[p] >= 0 && step_of[p] < t) ready.push_back(i);
        }
        std::stable_sort(ready.begin(), ready.end(), [&](int x, int y) { return height[x] > height[y]; });
        if (static_cast<int>(ready.size()) > U) ready.resize(U);
        std::vector<int> slot(U, -1);
        std::vector<char> placed(ready.size(), 0);
        // carries first: a picked joint whose parent sits in a track's registers continues that track
        for (size_t k = 0; k < ready.size(); ++k) {
            const int p = static_cast<int>(parents[ready[k]]);
            for (int u = 0; u < U; ++u)
                if (slot[u] < 0 && last[u] == p) {
                    slot[u] = ready[k], placed[k] = 2;
                    break;
                }
        }
        for (size_t k = 0; k < ready.size(); ++k) {
            if (placed[k]) continue;
            for (int u = 0; u < U; ++u)
                if (slot[u] < 0) {
                    slot[u] = ready[k], placed[k] = 1;
                    break;
                }
        }
        for (int u = 0; u < U; ++u) {
            uint32_t c = kTrackNoop;
            if (slot[u] >= 0) {
                const int i = slot[u], p = static_cast<int>(parents[i]);
                c = static_cast<uint32_t>(i) | (static_cast<uint32_t>(p) << 10) | (last[u] == p ? kTrackCarry : 0u);
                step_of[i] = t, ++done, ++w_done;
            }
            code[t * U + u] = c;
            last[u] = slot[u];  // a no-op breaks the carry
        }
        ++t;
    }
    if (chunk_first) chunk_first[(n_joints + W - 1) / W] = static_cast<uint16_t>(t);
    return t;
}

}  // namespace pmb
