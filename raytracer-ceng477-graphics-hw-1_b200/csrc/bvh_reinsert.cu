// bvh_reinsert.cu — insertion-based optimisation of a GPU-built BVH2 (reinsert_core.h has the algorithm and the
// reasons), as ONE cooperative launch: the builder's tree is unpacked into entity arrays, `rounds` rounds of
//   search + lock | grid barrier | verify | grid barrier | apply + refit | grid barrier
// run on persistent CTAs (one thread per entity and round), and the optimised tree is packed back over the builder's
// nodes when it is worth it (SAH cost below accept_ratio x the builder's, not deeper than the traversal stack).
// Nothing leaves the device; the host learns what happened from five fields of the BuildResult block.
//
// Determinism: the builder hands out node slots with an atomic counter, so slot numbers differ from run to run.  Locks
// are therefore keyed by a node's position in the depth-first order of the tree as built (computed here, once) — the
// numbering the host builder produces by construction — and never by its slot: equal gains resolve the same way in
// every run and exactly as in reinsert_optimize_host (tests compare the two).
//
// Input requirements (met by the top-down SAH builder, bvh_sah_device.cu): every node slot below *n_used is reachable
// from node 0, and no child is missing.  A tree that breaks them is left untouched.
#include <cfloat>
#include <cstdint>

#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include "build_device.h"
#include "reinsert_core.h"

namespace cg = cooperative_groups;

namespace rtb {

namespace {

constexpr int kRT = 256;
enum Counter : int { kBad = 0, kHeight = 1, kApplied = 2 };  // counters[kApplied + round]

// deterministic grid-wide sum: fixed-order tree inside the CTA, then every thread adds the CTA partials in order
__device__ float grid_sum(float v, float *partial, float *smem, cg::grid_group &grid) {
    const int t = threadIdx.x;
    smem[t] = v;
    __syncthreads();
    for (int off = kRT / 2; off > 0; off >>= 1) {
        if (t < off) smem[t] += smem[t + off];
        __syncthreads();
    }
    if (t == 0) partial[blockIdx.x] = smem[0];
    grid.sync();
    double s = 0;
    for (unsigned b = 0; b < gridDim.x; b++) s += (double) __ldcg(&partial[b]);
    grid.sync();  // partial[] may be reused
    return (float) s;
}

__global__ void __launch_bounds__(kRT) reinsert_kernel(DevTree t, ReinsertScratch s, int rounds, float accept_ratio, BuildResult *res) {
    __shared__ float smem[kRT];
    cg::grid_group grid = cg::this_grid();
    const int tid = blockIdx.x * kRT + threadIdx.x, nthreads = gridDim.x * kRT;
    if (*t.status != 0) return;  // uniform: the builder gave up
    const int n_used = *t.n_used, cap = t.cap, ne = s.ne;
    if (n_used < 3 || n_used > cap) return;
    const ReinsertView v = {s.box, s.left, s.right, s.parent};

    // ---- unpack: inner node i -> entity i, leaf range starting at primitive f -> entity cap + f -------------------
    for (int e = tid; e < ne; e += nthreads) {
        s.left[e] = s.right[e] = s.parent[e] = -1;
        s.lock[e] = 0ull;
        s.mv_y[e] = -1;
    }
    for (int i = tid; i < kApplied + kReinsertMaxRounds + 1; i += nthreads) s.counters[i] = 0;
    grid.sync();
    for (int i = tid; i < n_used; i += nthreads) {
        const HostNode nd = t.nodes[i];
        const int ch[2] = {nd.child0, nd.child1};
        const float *mns[2] = {nd.c0mn, nd.c1mn}, *mxs[2] = {nd.c0mx, nd.c1mx};
        int ce[2] = {-1, -1};
        bool bad = false;
        for (int c = 0; c < 2; c++) {
            if (ch[c] == kEmptyChild || ch[c] >= n_used || (ch[c] < 0 && ((~ch[c]) >> 3) + cap >= ne)) {
                bad = true;
                continue;
            }
            ce[c] = ch[c] >= 0 ? ch[c] : cap + ((~ch[c]) >> 3);
            if (ce[c] == 0 || atomicExch(&s.parent[ce[c]], i) != -1) bad = true;  // the root as a child / two parents
            if (ch[c] < 0) s.left[ce[c]] = ch[c];
            Aabb b;
            for (int k = 0; k < 3; k++) b.mn[k] = mns[c][k], b.mx[k] = mxs[c][k];
            s.box[ce[c]] = b;
        }
        if (bad) {
            atomicExch(&s.counters[kBad], 1);
        } else {
            s.left[i] = ce[0], s.right[i] = ce[1];
            if (i == 0) {
                Aabb a, b;
                for (int k = 0; k < 3; k++) a.mn[k] = nd.c0mn[k], a.mx[k] = nd.c0mx[k], b.mn[k] = nd.c1mn[k], b.mx[k] = nd.c1mx[k];
                s.box[0] = box_merge(a, b);
            }
        }
    }
    grid.sync();
    // every inner node but the root needs a parent (an orphan would be an unreachable node)
    for (int i = tid + 1; i < n_used; i += nthreads)
        if (__ldcg(&s.parent[i]) < 0) atomicExch(&s.counters[kBad], 1);
    grid.sync();
    if (__ldcg(&s.counters[kBad]) != 0) return;  // uniform

    // ---- canonical ids: depth-first (pre-order) index of every inner node, leaves keep cap + first primitive --------
    // inner nodes below each node, bottom-up: whoever completes a node (second arrival) carries on to its parent
    // (mv_y = arrivals, mv_pivot = subtree size: both are free until the first round)
    for (int i = tid; i < n_used; i += nthreads) s.mv_y[i] = 0;
    grid.sync();
    for (int i = tid; i < n_used; i += nthreads) {
        const int leaf_children = (s.left[s.left[i]] < 0 ? 1 : 0) + (s.left[s.right[i]] < 0 ? 1 : 0);
        if (leaf_children == 0) continue;
        if (leaf_children == 1 && atomicAdd(&s.mv_y[i], 1) != 1) continue;
        int cur = i;
        for (int guard = 0; guard < (1 << 24); guard++) {
            __threadfence();
            const int l = s.left[cur], r = s.right[cur];
            int count = 1;
            if (s.left[l] >= 0) count += __ldcg(&s.mv_pivot[l]);
            if (s.left[r] >= 0) count += __ldcg(&s.mv_pivot[r]);
            s.mv_pivot[cur] = count;
            __threadfence();
            cur = s.parent[cur];
            if (cur < 0) break;
            if (atomicAdd(&s.mv_y[cur], 1) != 1) break;  // the other subtree is not finished yet
        }
    }
    grid.sync();
    for (int e = tid; e < ne; e += nthreads) {
        int idx = e;
        if (e < n_used) {
            idx = 0;
            for (int cur = e, guard = 0; s.parent[cur] >= 0 && guard < kReinsertMaxWalk; guard++) {
                const int par = s.parent[cur];
                idx += 1;
                if (s.right[par] == cur && s.left[s.left[par]] >= 0) idx += __ldcg(&s.mv_pivot[s.left[par]]);
                cur = par;
            }
        }
        s.canon[e] = idx;
    }
    grid.sync();

    float my = 0;
    for (int i = tid; i < n_used; i += nthreads) my += reinsert_node_cost(v, i, kSahCostNode, kSahCostPrim);
    const float root_area = box_half_area(s.box[0]);
    const float sum_before = grid_sum(my, s.partial, smem, grid);
    const float cost_before = root_area > 0 ? kSahCostNode + sum_before / root_area : 0.0f;
    const float min_gain = 1e-6f * root_area;

    // ---- rounds --------------------------------------------------------------------------------------------------
    int moves = 0, done = 0;
    for (int round = 0; round < rounds; round++) {
        for (int x = tid; x < ne; x += nthreads) {
            ReinsertMove mv;
            int y = -1;
            if (reinsert_find(v, x, min_gain, mv)) {
                const unsigned long long key = reinsert_key(round, mv.gain, s.canon[x]);
                if (reinsert_paths(v, x, mv.y, mv.pivot, [&](int n) { atomicMax(&s.lock[n], key); return true; })) {
                    y = mv.y;
                    s.mv_pivot[x] = mv.pivot;
                    s.key[x] = key;
                }
            }
            s.mv_y[x] = y;
        }
        grid.sync();
        for (int x = tid; x < ne; x += nthreads) {
            const int y = s.mv_y[x];
            if (y < 0) continue;
            const unsigned long long key = s.key[x];
            if (!reinsert_paths(v, x, y, s.mv_pivot[x], [&](int n) { return __ldcg(&s.lock[n]) == key; })) s.mv_y[x] = -1;
        }
        grid.sync();
        int applied = 0;
        for (int x = tid; x < ne; x += nthreads) {
            const int y = s.mv_y[x];
            if (y < 0) continue;
            reinsert_apply(v, x, y, s.mv_pivot[x]);
            applied++;
        }
        if (applied) atomicAdd(&s.counters[kApplied + round], applied);
        grid.sync();
        const int n_applied = __ldcg(&s.counters[kApplied + round]);
        moves += n_applied;
        done = round + 1;
        if (n_applied == 0) break;  // uniform
    }

    // ---- verdict: cost and height of the optimised tree -----------------------------------------------------------
    my = 0;
    for (int i = tid; i < n_used; i += nthreads) {
        my += reinsert_node_cost(v, i, kSahCostNode, kSahCostPrim);
        int d = 1, a = s.parent[i];
        for (; a >= 0 && d < kReinsertMaxWalk; a = s.parent[a]) d++;
        atomicMax(&s.counters[kHeight], d);
    }
    const float sum_after = grid_sum(my, s.partial, smem, grid);
    const float cost_after = root_area > 0 ? kSahCostNode + sum_after / root_area : 0.0f;
    const int height = __ldcg(&s.counters[kHeight]);
    const bool accept = moves > 0 && cost_after < accept_ratio * cost_before && height <= kMaxTreeHeight;
    if (accept) {
        for (int i = tid; i < n_used; i += nthreads) {
            const int ce[2] = {s.left[i], s.right[i]};
            const Aabb b0 = s.box[ce[0]], b1 = s.box[ce[1]];
            HostNode nd;
            for (int k = 0; k < 3; k++) nd.c0mn[k] = b0.mn[k], nd.c0mx[k] = b0.mx[k], nd.c1mn[k] = b1.mn[k], nd.c1mx[k] = b1.mx[k];
            nd.child0 = ce[0] < cap ? ce[0] : s.left[ce[0]];
            nd.child1 = ce[1] < cap ? ce[1] : s.left[ce[1]];
            t.nodes[i] = nd;
        }
    }
    if (tid == 0) {
        res->reinsert_cost_before = cost_before;
        res->reinsert_cost_after = cost_after;
        res->reinsert_moves = moves;
        res->reinsert_rounds = done;
        res->reinsert_accepted = accept ? 1 : 0;
    }
}

}  // namespace

int reinsert_max_grid(int n_sms) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, reinsert_kernel, kRT, 0) != cudaSuccess || per_sm < 1) return 1;
    return n_sms * per_sm;
}

int reinsert_counter_slots() { return kApplied + kReinsertMaxRounds + 1; }

int enqueue_reinsert(const DevTree &t, const ReinsertScratch &s, int rounds, float accept_ratio, BuildResult *res, int grid,
                     cudaStream_t stream) {
    if (rounds <= 0 || s.ne >= kReinsertMaxEntities) return 0;
    if (rounds > kReinsertMaxRounds) rounds = kReinsertMaxRounds;
    long long want = ((long long) s.ne + kRT - 1) / kRT;
    if (want > grid) want = grid;
    if (want < 1) want = 1;
    DevTree tt = t;
    ReinsertScratch ss = s;
    void *args[] = {&tt, &ss, &rounds, &accept_ratio, &res};
    return (int) cudaLaunchCooperativeKernel((void *) reinsert_kernel, dim3((unsigned) want), dim3(kRT), args, 0, stream);
}

}  // namespace rtb
