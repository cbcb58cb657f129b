// Host-side compiler from a parents[] table to the per-joint program the chain
// kernels execute (fk, to_root_dual_quat).
//
// Joints are processed in index order (like ops/skeleton.py:51 and :234).  A
// joint whose parent is the joint right before it takes the parent transform
// from the registers of the thread walking the frame; any other joint fetches
// it from a shared-memory slot.  A joint is saved to a slot iff it has a child
// that does not come right after it; the slot is recycled after its last such
// child (linear-scan allocation), so the number of slots is the largest number
// of branch points alive at once -- 2 for the 22-joint body, 3 for a hand
// hanging off an arm off a spine -- not the joint count.
#pragma once
#include <stdint.h>

#include <vector>

#include "../../include/pymotion_b200.h"

namespace pmb {

struct ProgramInfo {
    int n_slots = 0;
    int status = PMB_OK;
    int bad_joint = -1;
};

// detach_root_children: to_root_dual_quat never composes with joint 0
// (ops/skeleton.py:236-237), so joint 0 is never saved and its children start
// fresh chains.
inline ProgramInfo build_joint_program(const int64_t *parents, int n_joints, bool detach_root_children,
                                       uint32_t *codes) {
    ProgramInfo info;
    if (n_joints < 1 || n_joints > PMB_MAX_JOINTS) {
        info.status = PMB_ERR_SHAPE;
        return info;
    }
    for (int i = 1; i < n_joints; ++i) {
        if (parents[i] < 0 || parents[i] >= i) {
            info.status = PMB_ERR_TOPOLOGY;
            info.bad_joint = i;
            return info;
        }
    }
    std::vector<int> last_use(n_joints, -1), slot_of(n_joints, -1);
    for (int i = 1; i < n_joints; ++i) {
        const int p = static_cast<int>(parents[i]);
        if (detach_root_children && p == 0) continue;
        if (p != i - 1) last_use[p] = i;  // ascending i: ends at the last non-adjacent child
    }
    std::vector<int> free_slots;  // kept sorted descending so back() is the lowest id
    int next_slot = 0;
    for (int i = 0; i < n_joints; ++i) {
        const int p = (i == 0) ? 0 : static_cast<int>(parents[i]);
        uint32_t src = 0xFFu, save = 0xFFu;
        const bool composes = i > 0 && !(detach_root_children && p == 0);
        if (composes && p != i - 1) {
            src = static_cast<uint32_t>(slot_of[p]);
            if (last_use[p] == i) {  // fetch happens before this joint's own save: the slot can be reused at once
                free_slots.push_back(slot_of[p]);
                for (size_t k = free_slots.size(); k > 1 && free_slots[k - 1] > free_slots[k - 2]; --k)
                    std::swap(free_slots[k - 1], free_slots[k - 2]);
            }
        }
        if (last_use[i] >= 0) {
            int s;
            if (!free_slots.empty()) {
                s = free_slots.back();
                free_slots.pop_back();
            } else {
                s = next_slot++;
            }
            if (s >= 0xFF) {
                info.status = PMB_ERR_TOPOLOGY;
                info.bad_joint = i;
                return info;
            }
            slot_of[i] = s;
            save = static_cast<uint32_t>(s);
        }
        codes[i] = src | (save << 8) | (static_cast<uint32_t>(p) << 16);
    }
    info.n_slots = next_slot;
    return info;
}

}  // namespace pmb
