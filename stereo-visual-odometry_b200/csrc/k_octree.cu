// k_octree.cu -- K3: ORBextractor::DistributeOctTree (ORBextractor.cpp:487-715) + ExtractorNode::DivideNode (:430-485).
//
// One CTA per (level, image).  The reference keeps a std::list of nodes, splits nodes into 4 children that are
// push_front-ed, and erases the parent.  Two facts make this data-parallel:
//   * nodes are only ever pushed to the front, so the list is always sorted by descending creation order; after a
//     round the list is   reverse(children in creation order) ++ (old list minus the split parents);
//   * which nodes are split in a round, and in which order, is a pure function of the node key counts:
//       round A (:556-615)  every node with > 1 key, in list order;
//       round B (:624-689)  nodes with > 1 key sorted by (count, address), walked from the back, stopping as soon as the
//                           list reaches N nodes.  Address ties are canonicalised to creation order (the reference's
//                           behaviour under a monotone allocator), i.e. descending count, then ascending list position.
//                           LIMITATION: the reference sorts pair<count, ExtractorNode*> (:636), so with an allocator that reuses
//                           freed blocks (glibc malloc after lNodes.erase) equal-count nodes can be split in a different order
//                           and 1-3 % of the keypoints selected near the N cut-off differ -- between two runs of the reference
//                           itself too (tests/test_reference_cpu.py prints the rate).  Bit-exactness against the compiled
//                           reference is therefore defined, and tested, under a bump allocator (oracle/_ref, monotone = true).
// So each round is: histogram keys into child quadrants -> (sort) -> prefix sums over parents -> rebuild the node
// array in list order -> relabel keys.  The node array lives in shared memory in list order; keys stay in global
// memory (L2-resident) with a 16-bit node position each.
//
// Output per level: for each final node, the key with the largest response (first in the reference's candidate order on
// ties, :696-712), in list order.
#include "svo_common.cuh"

namespace {

constexpr int OT = 256;

struct __align__(8) NodeBox {
    short x0, x1, y0, y1;
};

// in-place exclusive scan of a[0..n) in shared memory; returns the total.  All OT threads must call.
__device__ int block_scan_excl(int* a, int n, int* wtmp) {
    const int per = (n + OT - 1) / OT;
    const int b = threadIdx.x * per, e = min(b + per, n);
    int sum = 0;
    for (int i = b; i < e; ++i) sum += a[i];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) wtmp[warp] = incl;
    __syncthreads();
    int wbase = 0, total = 0;
#pragma unroll
    for (int w = 0; w < OT / 32; ++w) {
        const int v = wtmp[w];
        if (w < warp) wbase += v;
        total += v;
    }
    int run = wbase + incl - sum;
    for (int i = b; i < e; ++i) {
        const int v = a[i];
        a[i] = run;
        run += v;
    }
    __syncthreads();
    return total;
}

__device__ __forceinline__ int quadrant(const NodeBox& b, int x, int y) {
    const int xm = b.x0 + ((b.x1 - b.x0 + 1) >> 1);  // ceil(width / 2), DivideNode :432
    const int ym = b.y0 + ((b.y1 - b.y0 + 1) >> 1);
    return (x < xm ? 0 : 1) + (y < ym ? 0 : 2);  // n1, n2, n3, n4 (:464-474)
}

__global__ void __launch_bounds__(OT) k_octree(PyrGeom g, SlotMap map, const uint32_t* __restrict__ cand_all,
                                               const int* __restrict__ ncand, uint16_t* __restrict__ knode_all,
                                               uint32_t* __restrict__ lkp_all, int* __restrict__ nlkp, int* __restrict__ flags,
                                               uint32_t cand_per_slot, uint32_t lkp_per_slot, int ncap, int sortcap) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int l = blockIdx.x, slot = map.slot(blockIdx.y);
    const LevelGeom& G = g.lvl[l];
    const int n = ncand[slot * SVO_MAX_LEVELS + l];
    const uint32_t* __restrict__ cand = cand_all + (size_t)slot * cand_per_slot + G.cand_off;
    uint16_t* __restrict__ knode = knode_all + (size_t)slot * cand_per_slot + G.cand_off;
    uint32_t* __restrict__ out = lkp_all + (size_t)slot * lkp_per_slot + G.lkp_off;
    const int N = G.N;

    // shared arrays (ncap entries each unless noted)
    NodeBox* box[2];
    int* cnt[2];
    box[0] = reinterpret_cast<NodeBox*>(smem);
    box[1] = box[0] + ncap;
    cnt[0] = reinterpret_cast<int*>(box[1] + ncap);
    cnt[1] = cnt[0] + ncap;
    int* cc = cnt[1] + ncap;         // 4 * ncap: keys per child quadrant
    int* cbase = cc + 4 * ncap;      // creation index of the first child of a split parent, -1 if not split
    int* spos = cbase + ncap;        // new list position of a surviving node
    int* tmp = spos + ncap;          // scan scratch (ncap)
    uint32_t* skey = reinterpret_cast<uint32_t*>(tmp + ncap);  // sortcap: sort keys of round B
    int* wtmp = reinterpret_cast<int*>(skey + sortcap);        // 8 + 8 scalars
    int* sh = wtmp + 8;

    if (n == 0 || G.lkp_cap == 0) {
        if (threadIdx.x == 0) nlkp[slot * SVO_MAX_LEVELS + l] = 0;
        return;
    }

    // ---- roots (:491-545)
    int cur = 0;
    const int nIni = G.nIni;
    for (int i = threadIdx.x; i < nIni; i += OT) {
        NodeBox b;
        b.x0 = (short)(int)(__fmul_rn(G.hX, (float)i));
        b.x1 = (short)(int)(__fmul_rn(G.hX, (float)(i + 1)));
        b.y0 = 0;
        b.y1 = (short)G.Hp;
        box[0][i] = b;
        cnt[0][i] = 0;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < n; k += OT) {
        const int r = (int)__fdiv_rn((float)cand_x(cand[k]), G.hX);
        knode[k] = (uint16_t)r;
        atomicAdd(&cnt[0][r], 1);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nIni; i += OT) tmp[i] = cnt[0][i] > 0;
    __syncthreads();
    int size = block_scan_excl(tmp, nIni, wtmp);
    for (int i = threadIdx.x; i < nIni; i += OT)
        if (cnt[0][i] > 0) {
            box[1][tmp[i]] = box[0][i];
            cnt[1][tmp[i]] = cnt[0][i];
        }
    __syncthreads();
    for (int k = threadIdx.x; k < n; k += OT) knode[k] = (uint16_t)tmp[knode[k]];
    cur = 1;
    __syncthreads();

    bool modeB = false;
    for (int round = 0; round < 64; ++round) {
        NodeBox* B = box[cur];
        int* Cn = cnt[cur];
        NodeBox* Bn = box[cur ^ 1];
        int* Cnn = cnt[cur ^ 1];
        const int prevSize = size;

        // 1. histogram keys of multi-key nodes into child quadrants
        for (int i = threadIdx.x; i < 4 * size; i += OT) cc[i] = 0;
        __syncthreads();
        for (int k = threadIdx.x; k < n; k += OT) {
            const int i = knode[k];
            if (Cn[i] > 1) {
                const uint32_t c = cand[k];
                atomicAdd(&cc[4 * i + quadrant(B[i], cand_x(c), cand_y(c))], 1);
            }
        }
        __syncthreads();

        // 2. which parents are split, creation index of their first child
        int C;  // children created this round
        if (!modeB) {
            for (int i = threadIdx.x; i < size; i += OT)
                tmp[i] = Cn[i] > 1 ? (cc[4 * i] > 0) + (cc[4 * i + 1] > 0) + (cc[4 * i + 2] > 0) + (cc[4 * i + 3] > 0) : 0;
            __syncthreads();
            C = block_scan_excl(tmp, size, wtmp);
            for (int i = threadIdx.x; i < size; i += OT) cbase[i] = Cn[i] > 1 ? tmp[i] : -1;
            __syncthreads();
        } else {
            // sort parents by (count desc, list position asc): key = count << 12 | (4095 - pos), descending
            int M = 32;
            while (M < size) M <<= 1;
            for (int i = threadIdx.x; i < M; i += OT)
                skey[i] = (i < size && Cn[i] > 1) ? (((uint32_t)Cn[i] << 12) | (uint32_t)(4095 - i)) : 0u;
            __syncthreads();
            for (int k2 = 2; k2 <= M; k2 <<= 1)
                for (int j = k2 >> 1; j > 0; j >>= 1) {
                    for (int i = threadIdx.x; i < M; i += OT) {
                        const int p = i ^ j;
                        if (p > i) {
                            const uint32_t a = skey[i], b = skey[p];
                            const bool desc = (i & k2) == 0;
                            if (desc ? a < b : a > b) {
                                skey[i] = b;
                                skey[p] = a;
                            }
                        }
                    }
                    __syncthreads();
                }
            // processing order j -> node; net growth of the list per split = children - 1
            for (int j = threadIdx.x; j < size; j += OT) {
                const uint32_t key = skey[j];
                int v = 0;
                if (key) {
                    const int i = 4095 - (int)(key & 4095u);
                    v = (cc[4 * i] > 0) + (cc[4 * i + 1] > 0) + (cc[4 * i + 2] > 0) + (cc[4 * i + 3] > 0) - 1;
                }
                tmp[j] = v;
            }
            for (int i = threadIdx.x; i < size; i += OT) cbase[i] = -1;
            if (threadIdx.x == 0) sh[0] = size;  // number of parents processed before the break (default: all)
            __syncthreads();
            block_scan_excl(tmp, size, wtmp);  // tmp[j] = growth before processing j
            // the walk breaks after the first j whose split brings the list to >= N (:682-683)
            for (int j = threadIdx.x; j < size; j += OT) {
                const uint32_t key = skey[j];
                if (!key) continue;
                const int i = 4095 - (int)(key & 4095u);
                const int growth = (cc[4 * i] > 0) + (cc[4 * i + 1] > 0) + (cc[4 * i + 2] > 0) + (cc[4 * i + 3] > 0) - 1;
                if (size + tmp[j] + growth >= N) atomicMin(&sh[0], j);
            }
            __syncthreads();
            const int jbreak = sh[0];  // parents j <= jbreak are split
            // creation index: children of parent j start at (growth before j) + j   (each earlier parent adds growth+1)
            for (int j = threadIdx.x; j < size; j += OT) {
                const uint32_t key = skey[j];
                if (key && j <= jbreak) cbase[4095 - (int)(key & 4095u)] = tmp[j] + j;
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                // total children = creation index after the last split parent
                int last = min(jbreak, size - 1);
                while (last >= 0 && !skey[last]) --last;  // keys are sorted, zeros (non-parents) are at the end
                int total = 0;
                if (last >= 0) {
                    const int i = 4095 - (int)(skey[last] & 4095u);
                    total = tmp[last] + last + (cc[4 * i] > 0) + (cc[4 * i + 1] > 0) + (cc[4 * i + 2] > 0) + (cc[4 * i + 3] > 0);
                }
                sh[1] = total;
            }
            __syncthreads();
            C = sh[1];
        }

        // 3. survivors keep their relative order behind the new children
        for (int i = threadIdx.x; i < size; i += OT) tmp[i] = cbase[i] < 0;
        __syncthreads();
        const int nsurv = block_scan_excl(tmp, size, wtmp);
        const int newSize = C + nsurv;
        if (newSize > ncap) {  // cannot happen for lkp_cap = max(N + 3, 4 nIni) + 1; fail loudly if it does
            if (threadIdx.x == 0) {
                atomicOr(flags + slot, SVO_FLAG_NODE_OVERFLOW);
                nlkp[slot * SVO_MAX_LEVELS + l] = 0;
            }
            return;
        }
        for (int i = threadIdx.x; i < size; i += OT) {
            if (cbase[i] < 0) {
                spos[i] = C + tmp[i];
                Bn[C + tmp[i]] = B[i];
                Cnn[C + tmp[i]] = Cn[i];
            } else {
                const NodeBox b = B[i];
                const int xm = b.x0 + ((b.x1 - b.x0 + 1) >> 1), ym = b.y0 + ((b.y1 - b.y0 + 1) >> 1);
                int idx = cbase[i];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int kq = cc[4 * i + q];
                    if (kq == 0) continue;
                    NodeBox nb;
                    nb.x0 = (q & 1) ? (short)xm : b.x0;
                    nb.x1 = (q & 1) ? b.x1 : (short)xm;
                    nb.y0 = (q & 2) ? (short)ym : b.y0;
                    nb.y1 = (q & 2) ? b.y1 : (short)ym;
                    const int pos = C - 1 - idx;  // later children sit closer to the front
                    Bn[pos] = nb;
                    Cnn[pos] = kq;
                    ++idx;
                }
            }
        }
        __syncthreads();
        // 4. relabel keys
        for (int k = threadIdx.x; k < n; k += OT) {
            const int i = knode[k];
            if (cbase[i] < 0) {
                knode[k] = (uint16_t)spos[i];
            } else {
                const uint32_t c = cand[k];
                const int q = quadrant(B[i], cand_x(c), cand_y(c));
                int r = 0;
                for (int qq = 0; qq < q; ++qq) r += cc[4 * i + qq] > 0;
                knode[k] = (uint16_t)(C - 1 - (cbase[i] + r));
            }
        }
        // children with more than one key (nToExpand, :568-610)
        if (threadIdx.x == 0) sh[2] = 0;
        __syncthreads();
        if (!modeB) {
            int mine = 0;
            for (int i = threadIdx.x; i < C; i += OT) mine += Cnn[i] > 1;
            if (mine) atomicAdd(&sh[2], mine);
        }
        __syncthreads();
        const int nToExpand = sh[2];
        size = newSize;
        cur ^= 1;
        // 5. termination (:619-626, :686-687)
        if (size >= N || size == prevSize) break;
        if (!modeB && size + nToExpand * 3 > N) modeB = true;
    }

    // ---- best key per node (:696-712): highest response, earliest candidate on ties
    uint32_t* best = reinterpret_cast<uint32_t*>(cc);
    for (int i = threadIdx.x; i < size; i += OT) best[i] = 0;
    __syncthreads();
    for (int k = threadIdx.x; k < n; k += OT)
        atomicMax(&best[knode[k]], ((uint32_t)cand_r(cand[k]) << 24) | (0xFFFFFFu - (uint32_t)k));
    __syncthreads();
    if (size > G.lkp_cap) {
        if (threadIdx.x == 0) {
            atomicOr(flags + slot, SVO_FLAG_KP_OVERFLOW);
            nlkp[slot * SVO_MAX_LEVELS + l] = 0;
        }
        return;
    }
    for (int i = threadIdx.x; i < size; i += OT) out[i] = cand[0xFFFFFFu - (best[i] & 0xFFFFFFu)];
    if (threadIdx.x == 0) nlkp[slot * SVO_MAX_LEVELS + l] = size;
}

}  // namespace

int launch_octree(svo_ctx* c, const SlotMap& map, int n_img) {
    StageTimer t(c, 3);
    const PyrGeom& g = c->geom;
    const int ncap = c->octree_ncap;
    int sortcap = 32;
    while (sortcap < ncap) sortcap <<= 1;
    const size_t smem = (size_t)ncap * (2 * sizeof(NodeBox) + 2 * sizeof(int) + 4 * sizeof(int) + 3 * sizeof(int)) +
                        (size_t)sortcap * sizeof(uint32_t) + 16 * sizeof(int);
    SVO_CUDA_TRY(c, cudaFuncSetAttribute(k_octree, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(g.nlevels, n_img);
    k_octree<<<grid, OT, smem, c->stream>>>(g, map, c->d_cand, c->d_ncand, c->d_knode, c->d_lkp, c->d_nlkp, c->d_flags,
                                            c->cand_per_slot, c->lkp_per_slot, ncap, sortcap);
    ++c->launches;
    SVO_CUDA_TRY(c, cudaGetLastError());
    return SVO_OK;
}
