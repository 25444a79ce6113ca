// CLUSTER kernels: cluster_sv_signatures (SVIM_CLUSTER.py:7-26) for all six types in one pass.
//
//   k_sig_to_csig / k_cluster_keys / radix sorts   form_partitions' sorted(key=get_key)  (:19)
//   k_partition_heads + select                     the linear partition split           (:21-28)
//   host PyRandom                                  seed(1524) / sample(partition, 100)  (:129-134)
//   k_ins_pairs + k_myers_pairs                    haplotype edit distances             (:32-45)
//   k_linkage                                      same-read dedup, condensed span-position matrix,
//                                                  nn-chain average linkage, flat cut   (:141-175)
//   k_members / k_consolidate                      consolidate_clusters_*, scores       (:183-303)
//   final radix sort                               sorted(..., key=(contig, (end+start)/2)) (:381)
//
// One warp owns one partition (m <= 100 after sampling): records are bulk-copied into shared
// memory, the condensed FP64 matrix (<= 39.6 KB) lives there too, min-search and Lance-Williams
// update are warp-parallel, the O(m) dendrogram post-processing runs on lane 0.
#pragma once
#include <cub/cub.cuh>
#include <algorithm>
#include "ctx.cuh"
#include "cluster.cuh"
#include "pyrandom.h"
#include "myers.cu"

// multi-GPU hook (nccl.cu): allgatherv of the per-shard cluster records before the final ordering
static int cluster_exchange(svimgpu_ctx* ctx, uint32_t* n_clusters, uint32_t* n_members, int local_status);

// ---------------------------------------------------------------------------------------------
__global__ void k_sig_to_csig(const svim_sig* s, uint32_t n, const int32_t* rank, svim_csig* out) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const svim_sig a = s[k];
    svim_csig c; memset(&c, 0, sizeof(c));
    c.start = (double)a.start; c.end = (double)a.end; c.dpos = 0.0;
    c.contig_a = rank[a.contig1]; c.contig_b = -1;
    c.read_id = a.qname_id; c.seq_len = a.seq_len; c.seq_off = a.seq_off; c.type = a.type; c.copies = a.copies;
    if (a.type == SVIM_DUP_INT) { c.dpos = (double)a.pos; c.contig_b = rank[a.contig2]; }
    else if (a.type == SVIM_BND) {
        c.dpos = (double)a.pos; c.contig_b = rank[a.contig2];
        c.dirs = (uint8_t)(((a.flags & SVIM_F_DIR1_REV) ? 1 : 0) | ((a.flags & SVIM_F_DIR2_REV) ? 2 : 0));
    } else if (a.type == SVIM_INV) c.dirs = (uint8_t)((a.flags >> SVIM_F_INVDIR_SHIFT) & 7);
    out[k] = c;
}

// get_key per type: (type, contig, end) | INS (type, contig, start) | DUP_INT (type, dest, source, dest_start)
// | BND (type, contig1, pos1)     (SVSignature.py:21-23, 70-72, 132-135, 232-233)
__global__ void k_cluster_keys(const svim_csig* c, uint32_t n, uint64_t* coord, uint64_t* group, uint32_t* idx) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const svim_csig s = c[k];
    double key = s.end;
    if (s.type == SVIM_INS || s.type == SVIM_BND) key = s.start;
    else if (s.type == SVIM_DUP_INT) key = s.dpos;
    coord[k] = double_key(key);
    uint64_t k1 = (uint64_t)(uint32_t)s.contig_a, k2 = 0;
    if (s.type == SVIM_DUP_INT) { k1 = (uint64_t)(uint32_t)s.contig_b; k2 = (uint64_t)(uint32_t)s.contig_a; }
    group[k] = ((uint64_t)s.type << 60) | (k1 << 30) | k2;
    idx[k] = k;
}

template <class T>
__global__ void k_gather(const T* src, const uint32_t* order, uint32_t n, T* dst) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) dst[k] = src[order[k]];
}

__global__ void k_partition_heads(const svim_csig* s, const uint64_t* group, uint32_t n, double max_distance, uint8_t* head, uint32_t* type_start) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    bool h = true;
    const svim_csig b = s[k];
    if (k > 0) {
        const svim_csig a = s[k - 1];
        if (group[k] == group[k - 1]) h = gap_exceeds(b.type, a.start, a.end, a.dpos, b.start, b.dpos, max_distance);
        if (a.type != b.type) for (int t = a.type + 1; t <= b.type; ++t) type_start[t] = k;
    } else {
        for (int t = 0; t <= b.type; ++t) type_start[t] = 0;
    }
    if (k == n - 1) for (int t = b.type + 1; t <= 7; ++t) type_start[t] = n;
    head[k] = h ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------
// Partition plan on the device.  The host used to walk every partition (sizes, sample offsets, pair offsets, work lists,
// shard cuts): O(P + n) per step on every rank.  Now it sees a 256-word header and the sizes of the partitions above 100 —
// the only thing the sequential sampling stream (seed(1524) / sample(partition, 100), SVIM_clustering.py:129-134) consumes.
struct PMeta { unsigned long long pairs, cost; uint32_t m, large; };      // per partition; scanned with operator+
struct PMetaSum { __host__ __device__ __forceinline__ PMeta operator()(const PMeta& a, const PMeta& b) const {
    PMeta r; r.pairs = a.pairs + b.pairs; r.cost = a.cost + b.cost; r.m = a.m + b.m; r.large = a.large + b.large; return r; } };
struct LargePart { uint32_t p, size, type; };

enum { HDR_NLARGE = 0, HDR_NSAMP = 1, HDR_PAIRS = 2 /* 64-bit */, HDR_LO = 4, HDR_HI = 5, HDR_NSMALL = 6, HDR_NBIG = 7, HDR_NINS = 8, HDR_SLO = 9, HDR_SHI = 10 /* sample slots of this rank's partitions */,
       HDR_NPART_T = 16, HDR_NLARGE_T = 24, HDR_DUP_T = 32, HDR_NCL_T = 40, HDR_CUTS = 64, HDR_WORDS = 256, HDR_MAX_RANKS = 128,
       HDR_LARGE_INLINE = 2048 };

// sample size, pair count, cost estimate of every partition (+ a zero element at P so the exclusive scan ends on the totals)
__global__ void k_part_meta(const svim_csig* s, const uint32_t* part_off, uint32_t P, uint32_t n, PMeta* meta, uint8_t* ptype, uint32_t* hdr,
                            int32_t band_num, int32_t band_add) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    PMeta me; me.pairs = 0; me.cost = 0; me.m = 0; me.large = 0;
    int type = -1;
    if (p < P) {
        const uint32_t b = part_off[p], e = p + 1 < P ? part_off[p + 1] : n, sz = e - b;
        type = s[b].type;
        const uint32_t m = sz > 100 ? 100u : sz;
        me.m = m; me.large = sz > 100 ? 1u : 0u;
        ptype[p] = (uint8_t)type;
        // cost in ~lane-cycles: linkage ~ m^2; insertion pairs ~ columns x window blocks of the edit-distance kernel, from the
        // mean inserted length of (up to) the first 128 members and the spread of their start positions
        unsigned long long cost = 64ull + 50ull * m * m;
        if (type == SVIM_INS && m > 1) {
            me.pairs = (unsigned long long)m * (m - 1) / 2;
            const uint32_t cnt = sz > 128 ? 128u : sz;
            unsigned long long sum = 0;
            for (uint32_t k = 0; k < cnt; ++k) sum += s[b + k].seq_len;
            const double spread = s[b + cnt - 1].start - s[b].start;
            unsigned long long L = sum / cnt + (unsigned long long)(spread > 0 ? spread / 3.0 : 0.0);
            const unsigned long long band = band_num > 0 ? ((L * (unsigned long long)band_num) >> 10) + (unsigned long long)(band_add > 0 ? band_add : 0) + 63ull : L + 31ull;
            cost += me.pairs * (L * (band < L + 31ull ? band : L + 31ull)) / 57ull;
        }
        me.cost = cost;
    }
    if (p <= P) meta[p] = me;
    // per-type partition counts: partitions come grouped by type, so a warp rarely holds more than one
    const int t = type < 0 ? -1 : (type > 5 ? 5 : type);
    const unsigned grp = __match_any_sync(0xffffffffu, t);
    if (t >= 0) {
        const int lane = threadIdx.x & 31;
        if (lane == __ffs(grp) - 1) atomicAdd(hdr + HDR_NPART_T + t, (uint32_t)__popc(grp));
        const unsigned lg = __ballot_sync(grp, me.large != 0);
        if (lg && lane == __ffs(grp) - 1) atomicAdd(hdr + HDR_NLARGE_T + t, (uint32_t)__popc(lg));
    }
}

// shard cuts by cost (every rank computes the same cuts from the same scan), totals
__global__ void k_part_plan(const PMeta* pref, uint32_t P, int shard_rank, int shard_n, uint32_t* hdr) {
    const int r = threadIdx.x;
    __shared__ uint32_t cut[HDR_MAX_RANKS + 1];
    if (r <= shard_n) {
        uint32_t c = r == 0 ? 0u : P;
        if (r > 0 && r < shard_n) {
            const unsigned long long total = pref[P].cost;
            const unsigned long long target = (total / (unsigned)shard_n) * (unsigned)r + ((total % (unsigned)shard_n) * (unsigned)r) / (unsigned)shard_n;
            uint32_t lo = 0, hi = P;                       // first p with pref[p].cost >= target
            while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (pref[mid].cost < target) lo = mid + 1; else hi = mid; }
            c = lo;
        }
        cut[r] = c;
    }
    __syncthreads();
    if (r == 0) {
        for (int q = 1; q <= shard_n; ++q) if (cut[q] < cut[q - 1]) cut[q] = cut[q - 1];
        for (int q = 0; q <= shard_n; ++q) hdr[HDR_CUTS + q] = cut[q];
        hdr[HDR_LO] = cut[shard_rank]; hdr[HDR_HI] = cut[shard_rank + 1];
        hdr[HDR_SLO] = pref[cut[shard_rank]].m; hdr[HDR_SHI] = pref[cut[shard_rank + 1]].m;
        hdr[HDR_NLARGE] = pref[P].large; hdr[HDR_NSAMP] = pref[P].m;
        const unsigned long long pt = pref[P].pairs;
        hdr[HDR_PAIRS] = (uint32_t)pt; hdr[HDR_PAIRS + 1] = (uint32_t)(pt >> 32);
    }
}

// dense offset arrays for the kernels downstream, the list of partitions above 100 (all of them: the sampling stream is
// sequential per type), and this rank's three work lists  [linkage m <= 32 | linkage m > 32 | insertion pair lists]
__global__ void k_part_lists(const PMeta* meta, const PMeta* pref, const uint8_t* ptype, const uint32_t* part_off, uint32_t P, uint32_t n,
                             uint32_t* samp_off, uint64_t* pair_off, LargePart* large_list, uint32_t* plist, uint32_t* hdr) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    int which = -1; bool ins = false;
    if (p <= P) {
        const PMeta pr = pref[p];
        samp_off[p] = pr.m; pair_off[p] = pr.pairs;
        if (p < P) {
            const PMeta me = meta[p];
            if (me.large) { const uint32_t b = part_off[p], e = p + 1 < P ? part_off[p + 1] : n; large_list[pr.large] = LargePart{p, e - b, ptype[p]}; }
            if (p >= hdr[HDR_LO] && p < hdr[HDR_HI]) { which = me.m <= 32 ? 0 : 1; ins = me.pairs != 0; }
        }
    }
#pragma unroll
    for (int l = 0; l < 3; ++l) {
        const bool mine = l < 2 ? which == l : ins;
        const unsigned msk = __ballot_sync(0xffffffffu, mine);
        if (!msk) continue;
        uint32_t base = 0;
        if (lane == __ffs(msk) - 1) base = atomicAdd(hdr + HDR_NSMALL + l, (uint32_t)__popc(msk));
        base = __shfl_sync(0xffffffffu, base, __ffs(msk) - 1);
        if (mine) plist[(size_t)l * P + base + __popc(msk & ((1u << lane) - 1))] = p;
    }
}

// sample slots of this rank's partitions: all members in key order, or the host's 100 picks (sample order = member order, :172-175)
__global__ void __launch_bounds__(128) k_fill_samples(const uint32_t* part_off, const uint32_t* samp_off, const PMeta* meta, const PMeta* pref, uint32_t lo,
                                                       uint32_t hi, const int32_t* picks, uint32_t* samp_idx) {
    const int lane = threadIdx.x & 31;
    const uint32_t p = lo + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (p >= hi) return;
    const uint32_t b = part_off[p], s0 = samp_off[p], m = meta[p].m;
    if (meta[p].large) { const int32_t* pk = picks + (size_t)pref[p].large * 100; for (uint32_t k = lane; k < m; k += 32) samp_idx[s0 + k] = b + (uint32_t)pk[k]; }
    else for (uint32_t k = lane; k < m; k += 32) samp_idx[s0 + k] = b + k;
}

// ---- segmented INS blob (multi-GPU): the bytes of a signature's inserted sequence live on the rank that collected it -------------
// Each rank copies the sequences its own partitions will compare (the sample slots [s_lo, s_lo + S)) into one compact blob and
// points those records at it; everything downstream (pair lists, code image, edit-distance kernels) reads that blob as before.
struct SegTab { const int64_t* base; const uint8_t* const* ptr; int R; };

__device__ __forceinline__ int seg_rank_of(const SegTab& t, int64_t off) {
    int r = 0;
    while (r + 1 < t.R && off >= t.base[r + 1]) ++r;
    return r;
}

// slot bytes per sample: the sequence plus its source's misalignment, in whole 16-byte words (so the copy is aligned word for word)
__global__ void k_shard_ins_len(const svim_csig* sorted, const uint32_t* samp_idx, uint32_t s_lo, uint32_t S, SegTab t, uint64_t* len) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k > S) return;
    uint64_t v = 0;
    if (k < S) {
        const svim_csig c = sorted[samp_idx[s_lo + k]];
        if (c.type == SVIM_INS && c.seq_len > 0) {
            const int r = seg_rank_of(t, (int64_t)c.seq_off);
            const uint64_t mis = (uint64_t)((int64_t)c.seq_off - t.base[r]) & 15ull;
            v = (mis + (uint64_t)c.seq_len + 15ull) & ~15ull;
        }
    }
    len[k] = v;
}

// one warp per sample: 16-byte loads from the owner's blob (own HBM, or a peer's over NVLink), record re-pointed
__global__ void __launch_bounds__(256) k_shard_ins_copy(svim_csig* sorted, const uint32_t* samp_idx, uint32_t s_lo, uint32_t S, SegTab t,
                                                         const uint64_t* off, uint8_t* dst) {
    const int lane = threadIdx.x & 31;
    const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= S) return;
    const uint32_t idx = samp_idx[s_lo + w];
    const svim_csig c = sorted[idx];
    if (c.type != SVIM_INS || c.seq_len == 0) return;
    const int r = seg_rank_of(t, (int64_t)c.seq_off);
    const int64_t rel = (int64_t)c.seq_off - t.base[r];
    const int64_t mis = rel & 15;
    const uint4* src = (const uint4*)(t.ptr[r] + (rel - mis));
    uint4* d = (uint4*)(dst + off[w]);
    const uint32_t words = (uint32_t)((mis + (int64_t)c.seq_len + 15) >> 4);
    for (uint32_t q = lane; q < words; q += 32) d[q] = src[q];
    if (lane == 0) sorted[idx].seq_off = off[w] + (uint64_t)mis;
}

// ---------------------------------------------------------------------------------------------
// per-partition layout in shared memory (one warp per CTA)
struct PartSmem {
    double* start; double* end; double* dpos; uint32_t* read; uint8_t* dirs; uint8_t* dup; int* kidx;
    double* D; int* size; int* chain; int* ux; int* uy; double* ud;
    LinkScratch ls; int* T;
};

__host__ __device__ inline size_t part_smem_bytes(int M) {
    size_t dbl = 3 * (size_t)M + (size_t)M * (M - 1) / 2 + 3 * (size_t)M;      // start,end,dpos,D,ud,zd,md
    size_t i32 = 14 * (size_t)M + 16;                                          // read,kidx,size,chain,ux,uy,zx,zy,order,parent(2M),stack,T + slack
    size_t u8 = 4 * (size_t)M + 16;                                            // dirs,dup,visited(2M)
    return dbl * 8 + i32 * 4 + u8 + 64;
}

__device__ inline PartSmem carve(unsigned char* base, int M) {
    PartSmem p;
    double* d = (double*)base;
    p.start = d; d += M; p.end = d; d += M; p.dpos = d; d += M;
    p.D = d; d += (size_t)M * (M - 1) / 2;
    p.ud = d; d += M; p.ls.zd = d; d += M; p.ls.md = d; d += M;
    int* i = (int*)d;
    p.read = (uint32_t*)i; i += M; p.kidx = i; i += M; p.size = i; i += M; p.chain = i; i += M; p.ux = i; i += M; p.uy = i; i += M;
    p.ls.zx = i; i += M; p.ls.zy = i; i += M; p.ls.order = i; i += M; p.ls.parent = i; i += 2 * M; p.ls.stack = i; i += M; p.T = i; i += M;
    unsigned char* u = (unsigned char*)i;
    p.dirs = u; u += M; p.dup = u; u += M; p.ls.visited = u; u += 2 * M;
    return p;
}

struct LinkArgs {
    const svim_csig* sig;            // key-sorted signatures
    const uint32_t* samp_off;        // P+1
    const uint32_t* samp_idx;        // positions in `sig`
    const uint32_t* plist;           // partition ids handled by this launch
    uint32_t n_list;
    const uint64_t* pair_off;        // per partition: base into pair_ed (INS only)
    const int32_t* pair_ed;
    ClusterParams cp;
    int32_t* labels;                 // per sample slot: flat cluster id, 0 = same-read duplicate
    uint32_t* part_ncl; uint32_t* part_nkept;
    uint32_t* err;
    uint32_t* dup_t;                 // per type: sampled signatures dropped as same-read duplicates
};

// nn-chain average linkage on the condensed matrix D (mk points) -> unsorted merge rows
__device__ void nn_chain_warp(PartSmem& s, int mk, int lane) {
    for (int i = lane; i < mk; i += 32) s.size[i] = 1;
    __syncwarp();
    int chain_len = 0;
    for (int k = 0; k < mk - 1; ++k) {
        int x = 0, y = 0; double cur = 0.0;
        if (chain_len == 0) {
            // first i with size > 0
            int found = mk;
            for (int b = 0; b < mk; b += 32) {
                int i = b + lane;
                unsigned msk = __ballot_sync(FULL, i < mk && s.size[i] > 0);
                if (msk) { found = b + __ffs(msk) - 1; break; }
            }
            if (lane == 0) s.chain[0] = found;
            chain_len = 1;
            __syncwarp();
        }
        for (;;) {
            x = s.chain[chain_len - 1];
            int yprev = -1; double cprev = INFINITY;
            if (chain_len > 1) { yprev = s.chain[chain_len - 2]; cprev = s.D[cidx_any(mk, x, yprev)]; }
            // argmin over active i != x of D[x,i], lowest index on ties
            double bd = INFINITY; int bi = 0x7fffffff;
            for (int i = lane; i < mk; i += 32) {
                if (i == x || s.size[i] == 0) continue;
                double d = s.D[cidx_any(mk, x, i)];
                if (d < bd) { bd = d; bi = i; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                double od = __shfl_xor_sync(FULL, bd, o); int oi = __shfl_xor_sync(FULL, bi, o);
                if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
            }
            if (bd < cprev) { y = bi; cur = bd; } else { y = yprev; cur = cprev; }
            if (chain_len > 1 && y == yprev) break;
            if (lane == 0) s.chain[chain_len] = y;
            chain_len++;
            __syncwarp();
        }
        chain_len -= 2;
        if (x > y) { int t = x; x = y; y = t; }
        const int nx = s.size[x], ny = s.size[y];
        __syncwarp();
        if (lane == 0) { s.ux[k] = x; s.uy[k] = y; s.ud[k] = cur; s.size[x] = 0; s.size[y] = nx + ny; }
        __syncwarp();
        const double fx = (double)nx, fy = (double)ny, fs = (double)(nx + ny);
        for (int i = lane; i < mk; i += 32) {
            if (i == y || s.size[i] == 0) continue;
            const int iy = cidx_any(mk, i, y);
            s.D[iy] = (fx * s.D[cidx_any(mk, i, x)] + fy * s.D[iy]) / fs;
        }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(32) k_linkage(LinkArgs a, int M) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x;
    if (blockIdx.x >= a.n_list) return;
    const uint32_t p = a.plist[blockIdx.x];
    const uint32_t s0 = a.samp_off[p];
    const int m = (int)(a.samp_off[p + 1] - s0);
    PartSmem s = carve(smem_raw, M);
    const int type = a.sig[a.samp_idx[s0]].type;
    for (int i = lane; i < m; i += 32) {
        const svim_csig c = a.sig[a.samp_idx[s0 + i]];
        s.start[i] = c.start; s.end[i] = c.end; s.dpos[i] = c.dpos; s.read[i] = c.read_id; s.dirs[i] = c.dirs; s.dup[i] = 0;
    }
    __syncwarp();
    int err = 0;
    const int64_t npairs = (int64_t)m * (m - 1) / 2;
    const int32_t* ed = (type == SVIM_INS && a.pair_ed) ? a.pair_ed + a.pair_off[p] : nullptr;
    // ---- same-read duplicates (SVIM_clustering.py:141-151); none for INV ---------------------------
    if (type != SVIM_INV && m > 1) {
        int i = 0, rowstart = 0;   // walk condensed order
        for (int64_t q = lane; q < npairs; q += 32) {
            while (q >= rowstart + (m - 1 - i)) { rowstart += m - 1 - i; ++i; }
            const int j = i + 1 + (int)(q - rowstart);
            if (s.read[i] != s.read[j]) continue;
            SigView va{s.start[i], s.end[i], s.dpos[i], s.read[i], s.dirs[i]}, vb{s.start[j], s.end[j], s.dpos[j], s.read[j], s.dirs[j]};
            double e = ed ? (double)ed[q] : 0.0;
            double d = spd(type, va, vb, a.cp, e, &err);
            if (d <= a.cp.cluster_max_distance) s.dup[j] = 1;
        }
        __syncwarp();
    }
    // ---- compact survivors, in order -----------------------------------------------------------------
    int mk = 0;
    for (int b = 0; b < m; b += 32) {
        const int i = b + lane;
        const bool keep = i < m && !s.dup[i];
        const unsigned msk = __ballot_sync(FULL, keep);
        if (keep) s.kidx[mk + __popc(msk & ((1u << lane) - 1))] = i;
        mk += __popc(msk);
    }
    __syncwarp();
    for (int i = lane; i < m; i += 32) a.labels[s0 + i] = 0;
    __syncwarp();
    int ncl = 1;
    if (mk == 1) {
        if (lane == 0) a.labels[s0 + s.kidx[0]] = 1;
    } else {
        // ---- condensed distance matrix (:158-169) --------------------------------------------------
        const int64_t kp = (int64_t)mk * (mk - 1) / 2;
        int i = 0, rowstart = 0;
        for (int64_t q = lane; q < kp; q += 32) {
            while (q >= rowstart + (mk - 1 - i)) { rowstart += mk - 1 - i; ++i; }
            const int j = i + 1 + (int)(q - rowstart);
            const int oi = s.kidx[i], oj = s.kidx[j];
            double d;
            if (type != SVIM_INV && s.read[oi] == s.read[oj]) d = 99999.0;
            else {
                SigView va{s.start[oi], s.end[oi], s.dpos[oi], s.read[oi], s.dirs[oi]}, vb{s.start[oj], s.end[oj], s.dpos[oj], s.read[oj], s.dirs[oj]};
                double e = ed ? (double)ed[cidx(m, oi, oj)] : 0.0;
                d = spd(type, va, vb, a.cp, e, &err);
            }
            s.D[q] = d;
        }
        __syncwarp();
        nn_chain_warp(s, mk, lane);
        if (lane == 0) {
            ncl = fcluster_from_chain(mk, s.ls, s.ux, s.uy, s.ud, a.cp.cluster_max_distance, s.T);
            for (int k = 0; k < mk; ++k) a.labels[s0 + s.kidx[k]] = s.T[k];
        }
        ncl = __shfl_sync(FULL, ncl, 0);
    }
    if (lane == 0) { a.part_ncl[p] = (uint32_t)ncl; a.part_nkept[p] = (uint32_t)mk; if (m > mk && a.dup_t) atomicAdd(a.dup_t + (type > 5 ? 5 : type), (uint32_t)(m - mk)); }
    if (__any_sync(FULL, err)) { if (lane == 0) atomicExch(a.err, 1u); }
}

// INS partitions: list the pairs whose edit distance will be read (gate at :70 passes) and schedule them: list q < 10 =
// unbanded bin by pattern length, 10 + b = banded first pass of shape b (myers_band_bin).  mode 0 counts per list (and, per
// unbanded bin, the banded pairs that may be handed over to it), mode 1 fills `work` at the list cursors.
__global__ void __launch_bounds__(128) k_ins_pairs(const svim_csig* sig, const uint8_t* ins_blob, GenomeView g, const uint32_t* samp_off,
                                                    const uint32_t* samp_idx, const uint32_t* plist, uint32_t n_list, const uint64_t* pair_off,
                                                    ClusterParams cp, int mode, MyersWork* work, uint64_t* work_key, uint32_t* bin_cursor,
                                                    uint32_t* retry_cap, int32_t band_num, int32_t band_add, int32_t use_tpp, uint32_t* err) {
    const int lane = threadIdx.x & 31;
    const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= n_list) return;
    const uint32_t p = plist[w];
    const uint32_t s0 = samp_off[p];
    const int m = (int)(samp_off[p + 1] - s0);
    const int64_t npairs = (int64_t)m * (m - 1) / 2;
    int i = 0, rowstart = 0;
    for (int64_t q0 = 0; q0 < npairs; q0 += 32) {
        const int64_t q = q0 + lane;
        int bin = -1, rbin = -1; uint32_t pa = 0, pb = 0, cost = 0;
        if (q < npairs) {
            while (q >= rowstart + (m - 1 - i)) { rowstart += m - 1 - i; ++i; }
            const int j = i + 1 + (int)(q - rowstart);
            pa = samp_idx[s0 + i]; pb = samp_idx[s0 + j];
            const svim_csig a = sig[pa], b = sig[pb];
            SigView va{a.start, a.end, 0, 0, 0}, vb{b.start, b.end, 0, 0, 0};
            if (ins_gate_needs_ed(va, vb, cp)) {
                HapSource ha, hb;
                if (pair_haps(a, b, ins_blob, g, ha, hb)) {
                    const int64_t la = ha.l1 + ha.l2 + ha.l3, lb = hb.l1 + hb.l2 + hb.l3;
                    const int64_t lm = la > lb ? la : lb, ln = la > lb ? lb : la;
                    const TppPlan tp = tpp_plan(lm, ln, band_num, band_add);
                    const int tq = use_tpp ? tpp_bucket_of(tp.B) : -1;
                    const int band = tq >= 0 ? -1 : myers_band_bin(lm, ln, band_num, band_add);
                    if (tq >= 0) {                                   // thread-per-pair window (banded or whole pattern)
                        bin = 2 * MYERS_BINS + tq; rbin = myers_bin_of(lm);
                        cost = (uint32_t)ln;                         // columns
                    } else if (band >= 0) {
                        bin = MYERS_BINS + band; rbin = myers_bin_of(lm);
                        cost = (uint32_t)(ln + (lm >> 6));          // ~ wavefront steps of the banded pass
                    } else {
                        bin = myers_bin_of(lm);
                        const uint64_t cells = (uint64_t)la * (uint64_t)lb;
                        cost = cells >> 6 > 0xffffffffull ? 0xffffffffu : (uint32_t)(cells >> 6);
                    }
                } else { atomicExch(err, 1u); }
            }
        }
        // lanes of the same list reserve their slots with one atomic (match.any groups them)
        {
            const unsigned grp = __match_any_sync(FULL, bin);
            if (bin >= 0) {
                const int leader = __ffs(grp) - 1;
                uint32_t base = 0;
                if (lane == leader) base = atomicAdd(bin_cursor + bin, (uint32_t)__popc(grp));
                base = __shfl_sync(grp, base, leader);
                if (mode == 1) {
                    const uint32_t at = base + __popc(grp & ((1u << lane) - 1));
                    MyersWork wk{pa, pb, (uint32_t)(pair_off[p] + q), 0}; work[at] = wk;
                    work_key[at] = ((uint64_t)bin << 32) | (uint64_t)(0xffffffffu - cost);   // longest pairs first inside a list (LPT)
                }
            }
        }
        if (mode == 0 && __any_sync(FULL, rbin >= 0)) {
#pragma unroll
            for (int bb = 0; bb < MYERS_BINS; ++bb) {
                const unsigned msk = __ballot_sync(FULL, rbin == bb);
                if (msk && lane == (__ffs(msk) - 1)) atomicAdd(retry_cap + bb, (uint32_t)__popc(msk));
            }
        }
    }
}

// members of every flat cluster, in sample order (SVIM_clustering.py:172-175)
__global__ void __launch_bounds__(128) k_members(const uint32_t* samp_off, const uint32_t* samp_idx, const uint32_t* order, const int32_t* labels,
                                                  const uint32_t* part_ncl, const uint32_t* cl_off, const uint32_t* mem_off, uint32_t P,
                                                  svim_cluster* clusters, uint32_t* members) {
    const int lane = threadIdx.x & 31;
    const uint32_t p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (p >= P) return;
    const uint32_t s0 = samp_off[p];
    const int m = (int)(samp_off[p + 1] - s0);
    const int ncl = (int)part_ncl[p];
    uint32_t out = mem_off[p];
    for (int c = 1; c <= ncl; ++c) {
        const uint32_t begin = out;
        for (int b = 0; b < m; b += 32) {
            const int i = b + lane;
            const bool hit = i < m && labels[s0 + i] == c;
            const unsigned msk = __ballot_sync(FULL, hit);
            if (hit) members[out + __popc(msk & ((1u << lane) - 1))] = order[samp_idx[s0 + i]];
            out += __popc(msk);
        }
        if (lane == 0) { svim_cluster& cl = clusters[cl_off[p] + c - 1]; cl.member_off = begin; cl.size = out - begin; }
    }
}

// consolidate_clusters_unilocal / _bilocal + calculate_score, one thread per cluster
__global__ void k_consolidate(const svim_csig* sig /* emission order */, const uint32_t* members, uint32_t n_clusters, svim_cluster* clusters,
                              uint32_t* err) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_clusters) return;
    svim_cluster cl = clusters[c];
    const uint32_t* mem = members + cl.member_off;
    const int n = (int)cl.size;
    const svim_csig first = sig[mem[0]];
    const int type = first.type;
    double sum_s = 0.0, sum_e = 0.0, sum_ds = 0.0, sum_de = 0.0;
    int left = 0, right = 0, all = 0, max_copies = 0;
    uint32_t dirs_or = 0, dirs_and = 3;
    double v[3 * 100];   // span, pos, dest pos   (n <= 100)
    for (int i = 0; i < n; ++i) {
        const svim_csig m = sig[mem[i]];
        sum_s += m.start; sum_e += m.end;
        v[3 * i] = m.end - m.start; v[3 * i + 1] = (m.end + m.start) / 2.0;
        double ds = 0.0, de = 0.0;
        if (type == SVIM_DUP_INT) { ds = m.dpos; de = m.dpos + (m.end - m.start); }
        else if (type == SVIM_DUP_INT_CAND) { ds = m.dpos; memcpy(&de, &m.seq_off, 8); }   // explicit destination end
        else if (type == SVIM_BND) { ds = m.dpos; de = m.dpos + 1.0; }
        sum_ds += ds; sum_de += de;
        v[3 * i + 2] = (de + ds) / 2.0;
        if (type == SVIM_INV) { left += m.dirs < 2; right += (m.dirs == 2 || m.dirs == 3); all += m.dirs == 4; }
        if (m.copies > max_copies) max_copies = m.copies;
        dirs_or |= m.dirs; dirs_and &= m.dirs;
    }
    const double a_s = sum_s / (double)n, a_e = sum_e / (double)n;
    const int has = n > 1;
    double sd_span = 0.0, sd_pos = 0.0;
    if (has) { sd_span = stdev_values(v, n, 3); sd_pos = stdev_values(v + 1, n, 3); }
    cl.type = (uint8_t)type;
    cl.start = py_round_int(a_s); cl.end = py_round_int(a_e);
    cl.dest_start = 0; cl.dest_end = 0; cl.dir1_rev = 0; cl.dir2_rev = 0;
    const double nanv = nan("");
    int n_eff = n;
    if (type == SVIM_INV) n_eff = (left < right ? left : right) + all;
    double span = a_e - a_s, ss = sd_span, sp = sd_pos;
    if (type == SVIM_DUP_TAN) {
        cl.dest_start = cl.end; cl.dest_end = cl.end + (int64_t)max_copies * (cl.end - cl.start);
    } else if (type == SVIM_DUP_INT || type == SVIM_DUP_INT_CAND) {
        const double d_s = sum_ds / (double)n, d_e = sum_de / (double)n;
        cl.dest_start = py_round_int(d_s); cl.dest_end = py_round_int(d_e);
        span = ((a_e - a_s) + (d_e - d_s)) / 2.0;                 // mean([..]) of two floats
        if (has) {
            // destination span == source span member-wise; destination position has its own spread
            const double dd_span = sd_span;   // stdev of (dest_end - dest_start) = stdev of source spans
            const double dd_pos = stdev_values(v + 2, n, 3);
            ss = (sd_span + dd_span) / 2.0; sp = (sd_pos + dd_pos) / 2.0;
        }
    } else if (type == SVIM_BND) {
        const double d_s = sum_ds / (double)n, d_e = sum_de / (double)n;
        cl.dest_start = py_round_int(d_s); cl.dest_end = py_round_int(d_e);
        if (dirs_or != dirs_and) atomicExch(err, 3u);             // assert len(directions) == 1
        cl.dir1_rev = first.dirs & 1; cl.dir2_rev = (first.dirs >> 1) & 1;
        span = 500.0;
        if (has) { ss = sd_pos; sp = stdev_values(v + 2, n, 3); }
    }
    if (has && span == 0.0 && type != SVIM_DUP_INT_CAND) atomicExch(err, 1u);   // ZeroDivisionError in calculate_score (candidates are scored on the host)
    cl.score = cluster_score(n_eff, has, ss, sp, span);
    cl.std_span = has ? ss : nanv; cl.std_pos = has ? sp : nanv;
    clusters[c] = cl;
}

// final list order: unilocal types sorted by (contig, (end+start)/2), bilocal keep partition order
__global__ void k_final_keys(const svim_cluster* cl, const svim_csig* sig, const uint32_t* members, uint32_t n, uint64_t* k1, uint64_t* k2, uint32_t* idx, uint32_t* ncl_t) {
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    const int ty = c < n ? (cl[c].type > 5 ? 5 : cl[c].type) : -1;
    const unsigned grp = __match_any_sync(0xffffffffu, ty);
    if (ty >= 0 && (int)(threadIdx.x & 31) == __ffs(grp) - 1) atomicAdd(ncl_t + ty, (uint32_t)__popc(grp));
    if (c >= n) return;
    const svim_cluster x = cl[c];
    const bool uni = x.type == SVIM_DEL || x.type == SVIM_INS || x.type == SVIM_INV;
    k1[c] = uni ? double_key((double)(x.end + x.start) / 2.0) : 0ull;
    const uint64_t contig = uni ? (uint64_t)(uint32_t)sig[members[x.member_off]].contig_a : 0ull;
    k2[c] = ((uint64_t)x.type << 60) | contig;
    idx[c] = c;
}

// ---------------------------------------------------------------------------------------------
static int sort_pairs_u64(svimgpu_ctx* ctx, DevBuf* keys, DevBuf* vals, uint32_t n, int end_bit, uint64_t** out_k, uint32_t** out_v) {
    cub::DoubleBuffer<uint64_t> dk(keys[0].as<uint64_t>(), keys[1].as<uint64_t>());
    cub::DoubleBuffer<uint32_t> dv(vals[0].as<uint32_t>(), vals[1].as<uint32_t>());
    size_t tmp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp, dk, dv, (int)n, 0, end_bit, ctx->stream);
    SVIM_CUDA(ctx->d_sort_tmp.ensure(tmp));
    SVIM_CUDA(cub::DeviceRadixSort::SortPairs(ctx->d_sort_tmp.p, tmp, dk, dv, (int)n, 0, end_bit, ctx->stream));
    *out_k = dk.Current(); *out_v = dv.Current();
    if (dk.Current() != keys[0].as<uint64_t>()) std::swap(keys[0], keys[1]);
    if (dv.Current() != vals[0].as<uint32_t>()) std::swap(vals[0], vals[1]);
    return 0;
}

static int cluster_run(svimgpu_ctx* ctx, svim_cluster_stats* stats, int shard_rank, int shard_n, bool partition_only = false) {
    if (!ctx->have_csig) {
        ctx->set_error(SVIMGPU_ERR_STATE, "no signatures selected for clustering");
        if (shard_n > 1) { uint32_t z0 = 0, z1 = 0; cluster_exchange(ctx, &z0, &z1, SVIMGPU_ERR_STATE); }      // the other ranks are on their way into the exchange
        return SVIMGPU_ERR_STATE;
    }
    cudaStream_t st = ctx->stream;
    const uint32_t n = (uint32_t)ctx->n_csig;
    svim_cluster_stats& cs = ctx->clstats;
    memset(&cs, 0, sizeof(cs));
    ctx->n_clusters_host = 0; ctx->n_members_host = 0; ctx->n_partitions = 0;
    ctx->clustered = true;
    if (n == 0) { if (stats) *stats = cs; return 0; }
    ClusterParams cp{ctx->params.partition_max_distance, ctx->params.position_distance_normalizer, ctx->params.edit_distance_normalizer,
                     ctx->params.cluster_max_distance};
    uint32_t n_clusters = 0, n_members = 0;
    bool stop_after_partition = false;
    // Everything up to the consolidated clusters of this rank's partitions.  In sharded mode a failure here must not return before the
    // other ranks have been told (they would wait in the exchange for ever): the status travels with the exchange's count all-gather.
    auto local = [&]() -> int {
    if (const char* f = getenv("SVIM_TEST_FAIL_RANK"))            // test hook (tests/multi_gpu_check.py): this rank fails before the exchange
        if (shard_n > 1 && atoi(f) == shard_rank) { ctx->set_error(SVIMGPU_ERR_DATA, "cluster: injected failure on rank %d", shard_rank); return SVIMGPU_ERR_DATA; }
    const uint32_t nb = (n + 255) / 256;
    // ---- sort by get_key -------------------------------------------------------------------------------
    uint64_t* gkey_sorted = nullptr; uint32_t* order = nullptr;
    {
        StageTimer t(ctx, T_KEYSORT);
        for (int k = 0; k < 2; ++k) { SVIM_CUDA(ctx->d_ckeys[k].ensure((size_t)n * 8)); SVIM_CUDA(ctx->d_cvals[k].ensure((size_t)n * 4)); }
        SVIM_CUDA(ctx->d_keys[0].ensure((size_t)n * 8)); SVIM_CUDA(ctx->d_keys[1].ensure((size_t)n * 8));
        uint64_t* group = ctx->d_keys[0].as<uint64_t>();
        { ctx->launches++; k_cluster_keys<<<nb, 256, 0, st>>>(ctx->d_csig.as<svim_csig>(), n, ctx->d_ckeys[0].as<uint64_t>(), group, ctx->d_cvals[0].as<uint32_t>()); }
        uint64_t* k_out; uint32_t* v_out;
        int rc = sort_pairs_u64(ctx, ctx->d_ckeys, ctx->d_cvals, n, 64, &k_out, &v_out); if (rc) return rc;
        { ctx->launches++; k_gather<uint64_t><<<nb, 256, 0, st>>>(group, v_out, n, ctx->d_ckeys[1].as<uint64_t>()); }
        std::swap(ctx->d_ckeys[0], ctx->d_ckeys[1]);
        rc = sort_pairs_u64(ctx, ctx->d_ckeys, ctx->d_cvals, n, 64, &k_out, &v_out); if (rc) return rc;
        gkey_sorted = k_out; order = v_out;
        SVIM_CUDA(ctx->d_csig_sorted.ensure((size_t)n * sizeof(svim_csig)));
        { ctx->launches++; k_gather<svim_csig><<<nb, 256, 0, st>>>(ctx->d_csig.as<svim_csig>(), order, n, ctx->d_csig_sorted.as<svim_csig>()); }
        SVIM_CUDA(ctx->d_order.ensure((size_t)n * 4));
        SVIM_CUDA(cudaMemcpyAsync(ctx->d_order.p, order, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
    }
    const svim_csig* sorted = ctx->d_csig_sorted.as<svim_csig>();
    // ---- partitions --------------------------------------------------------------------------------------
    uint32_t P = 0;
    {
        StageTimer t(ctx, T_PARTITION);
        SVIM_CUDA(ctx->d_head.ensure(n + 64)); SVIM_CUDA(ctx->d_part_off.ensure((size_t)(n + 2) * 4)); SVIM_CUDA(ctx->d_part_stats.ensure(64 * 4));
        uint32_t* d_ts = ctx->d_part_stats.as<uint32_t>();   // [0..6] type_start, [8] num selected, [9] err, [10] n_work, [12..13] cells
        SVIM_CUDA(cudaMemsetAsync(d_ts, 0, 64 * 4, st));
        { ctx->launches++; k_partition_heads<<<nb, 256, 0, st>>>(sorted, gkey_sorted, n, cp.partition_max_distance, ctx->d_head.as<uint8_t>(), d_ts); }
        size_t tmp = 0;
        cub::CountingInputIterator<uint32_t> cnt_it(0);
        cub::DeviceSelect::Flagged(nullptr, tmp, cnt_it, ctx->d_head.as<uint8_t>(), ctx->d_part_off.as<uint32_t>(), d_ts + 8, (int)n, st);
        SVIM_CUDA(ctx->d_sort_tmp.ensure(tmp));
        SVIM_CUDA(cub::DeviceSelect::Flagged(ctx->d_sort_tmp.p, tmp, cnt_it, ctx->d_head.as<uint8_t>(), ctx->d_part_off.as<uint32_t>(), d_ts + 8, (int)n, st));
        SVIM_CUDA(ctx->h_hdr.ensure((HDR_WORDS + 3 * HDR_LARGE_INLINE) * 4));
        uint32_t* h = ctx->h_hdr.as<uint32_t>();
        SVIM_CUDA(cudaMemcpyAsync(h, d_ts + 8, 4, cudaMemcpyDeviceToHost, st));
        SVIM_CUDA(cudaStreamSynchronize(st));
        P = h[0];
        ctx->n_partitions = P;
    }
    if (partition_only) { stop_after_partition = true; return 0; }
    if (shard_n > HDR_MAX_RANKS) { ctx->set_error(SVIMGPU_ERR_LIMIT, "more than %d ranks", (int)HDR_MAX_RANKS); return SVIMGPU_ERR_LIMIT; }
    // ---- plan on the device: sample sizes, pair offsets, shard cuts, work lists; host: the sampling stream ----------------
    uint32_t n_samp = 0, n_small = 0, n_big = 0, n_ins = 0, lo = 0, hi = P;
    uint64_t pair_total = 0;
    const int max_m_small = 32;
    {
        StageTimer t(ctx, T_SAMPLE);
        const uint32_t pb = (P + 1 + 255) / 256;
        SVIM_CUDA(ctx->d_pmeta.ensure((size_t)(P + 1) * sizeof(PMeta))); SVIM_CUDA(ctx->d_ppref.ensure((size_t)(P + 1) * sizeof(PMeta)));
        SVIM_CUDA(ctx->d_ptype.ensure(P + 16)); SVIM_CUDA(ctx->d_hdr.ensure(HDR_WORDS * 4));
        SVIM_CUDA(ctx->d_samp_off.ensure((size_t)(P + 1) * 4)); SVIM_CUDA(ctx->d_pair_off.ensure((size_t)(P + 1) * 8));
        SVIM_CUDA(ctx->d_plist.ensure((size_t)3 * P * 4 + 16)); SVIM_CUDA(ctx->d_samp_idx.ensure((size_t)n * 4 + 4));
        uint32_t* d_hdr = ctx->d_hdr.as<uint32_t>();
        PMeta* meta = ctx->d_pmeta.as<PMeta>(); PMeta* pref = ctx->d_ppref.as<PMeta>();
        SVIM_CUDA(cudaMemsetAsync(d_hdr, 0, HDR_WORDS * 4, st));
        { ctx->launches++; k_part_meta<<<pb, 256, 0, st>>>(sorted, ctx->d_part_off.as<uint32_t>(), P, n, meta, ctx->d_ptype.as<uint8_t>(), d_hdr,
                                                            ctx->myers_band_num, ctx->myers_band_add); }
        size_t tmp = 0;
        PMeta zero; zero.pairs = 0; zero.cost = 0; zero.m = 0; zero.large = 0;
        cub::DeviceScan::ExclusiveScan(nullptr, tmp, meta, pref, PMetaSum(), zero, (int)P + 1, st);
        SVIM_CUDA(ctx->d_sort_tmp.ensure(tmp));
        SVIM_CUDA(cub::DeviceScan::ExclusiveScan(ctx->d_sort_tmp.p, tmp, meta, pref, PMetaSum(), zero, (int)P + 1, st));
        { ctx->launches++; k_part_plan<<<1, HDR_MAX_RANKS + 1, 0, st>>>(pref, P, shard_n > 1 ? shard_rank : 0, shard_n > 1 ? shard_n : 1, d_hdr); }
        // the list of partitions above 100 cannot be longer than n / 101
        SVIM_CUDA(ctx->d_large_list.ensure(((size_t)n / 101 + 2) * sizeof(LargePart)));
        { ctx->launches++; k_part_lists<<<pb, 256, 0, st>>>(meta, pref, ctx->d_ptype.as<uint8_t>(), ctx->d_part_off.as<uint32_t>(), P, n, ctx->d_samp_off.as<uint32_t>(),
                                                             ctx->d_pair_off.as<uint64_t>(), ctx->d_large_list.as<LargePart>(), ctx->d_plist.as<uint32_t>(), d_hdr); }
        uint32_t* h = ctx->h_hdr.as<uint32_t>();
        SVIM_CUDA(cudaMemcpyAsync(h, d_hdr, HDR_WORDS * 4, cudaMemcpyDeviceToHost, st));
        const size_t inline_cap = std::min<size_t>(HDR_LARGE_INLINE, (size_t)n / 101 + 1);
        SVIM_CUDA(cudaMemcpyAsync(h + HDR_WORDS, ctx->d_large_list.p, inline_cap * sizeof(LargePart), cudaMemcpyDeviceToHost, st));
        SVIM_CUDA(cudaStreamSynchronize(st));
        const uint32_t n_large = h[HDR_NLARGE];
        n_samp = h[HDR_NSAMP]; pair_total = ((uint64_t)h[HDR_PAIRS + 1] << 32) | h[HDR_PAIRS];
        lo = h[HDR_LO]; hi = h[HDR_HI]; n_small = h[HDR_NSMALL]; n_big = h[HDR_NBIG]; n_ins = h[HDR_NINS];
        for (int ty = 0; ty < 6; ++ty) { cs.n_partitions[ty] = h[HDR_NPART_T + ty]; cs.large_partitions[ty] = h[HDR_NLARGE_T + ty]; }
        ctx->shard_lo = lo; ctx->shard_hi = hi;
        if (n_large) {
            SVIM_CUDA(ctx->h_picks.ensure((size_t)n_large * (100 * 4 + sizeof(LargePart))));
            int32_t* picks = ctx->h_picks.as<int32_t>();
            LargePart* lp = (LargePart*)(picks + (size_t)n_large * 100);
            const uint32_t got = (uint32_t)std::min<size_t>(n_large, inline_cap);
            memcpy(lp, h + HDR_WORDS, (size_t)got * sizeof(LargePart));
            if (n_large > got) {
                SVIM_CUDA(cudaMemcpyAsync(lp + got, ctx->d_large_list.as<LargePart>() + got, (size_t)(n_large - got) * sizeof(LargePart), cudaMemcpyDeviceToHost, st));
                SVIM_CUDA(cudaStreamSynchronize(st));
            }
            PyRandom rng; int cur_type = -1;                 // seed(1524) once per type (:129), one sample() per partition above 100 (:133)
            for (uint32_t k = 0; k < n_large; ++k) {
                if ((int)lp[k].type != cur_type) { rng.seed_int(1524); cur_type = (int)lp[k].type; }
                rng.sample100(lp[k].size, picks + (size_t)k * 100);
            }
            SVIM_CUDA(ctx->d_picks.ensure((size_t)n_large * 400));
            SVIM_CUDA(cudaMemcpyAsync(ctx->d_picks.p, picks, (size_t)n_large * 400, cudaMemcpyHostToDevice, st));
        }
        if (hi > lo) { ctx->launches++; k_fill_samples<<<(uint32_t)(((uint64_t)(hi - lo) * 32 + 127) / 128), 128, 0, st>>>(ctx->d_part_off.as<uint32_t>(), ctx->d_samp_off.as<uint32_t>(), meta, pref, lo, hi,
                                                                                            ctx->d_picks.as<int32_t>(), ctx->d_samp_idx.as<uint32_t>()); }
    }
    // ---- the inserted sequences this rank's partitions compare ------------------------------------------------------------
    const uint8_t* ins_blob = ctx->cluster_ins; int64_t ins_blob_bytes = ctx->cluster_ins_bytes;
    if (ctx->cluster_seg >= 0) {
        const SigSet& set = ctx->sets[ctx->cluster_seg];
        if (!set.segmented) { ctx->set_error(SVIMGPU_ERR_STATE, "cluster: the selected list was collected again since svimgpu_use_collected"); return SVIMGPU_ERR_STATE; }
        if (ctx->peer_failed) { ctx->set_error(SVIMGPU_ERR_NCCL, "cluster: a peer rank's insertion bytes could not be mapped (CUDA IPC); run with SVIM_PEER_INS=0"); return SVIMGPU_ERR_NCCL; }
        const uint32_t* h = ctx->h_hdr.as<uint32_t>();
        const uint32_t s_lo = h[HDR_SLO], S = h[HDR_SHI] - h[HDR_SLO];
        const int R = (int)set.seg_ptr.size();
        ins_blob = nullptr; ins_blob_bytes = 0;
        if (S > 0 && n_ins > 0) {
            StageTimer t(ctx, T_PEER_INS);
            SVIM_CUDA(ctx->d_seg_tab.ensure((size_t)(2 * R + 1) * 8));
            SVIM_CUDA(cudaMemcpyAsync(ctx->d_seg_tab.p, set.seg_base.data(), (size_t)(R + 1) * 8, cudaMemcpyHostToDevice, st));
            SVIM_CUDA(cudaMemcpyAsync(ctx->d_seg_tab.as<int64_t>() + R + 1, set.seg_ptr.data(), (size_t)R * 8, cudaMemcpyHostToDevice, st));
            SegTab tab{ctx->d_seg_tab.as<int64_t>(), (const uint8_t* const*)(ctx->d_seg_tab.as<int64_t>() + R + 1), R};
            SVIM_CUDA(ctx->d_shard_len.ensure((size_t)(S + 1) * 16));
            uint64_t* len = ctx->d_shard_len.as<uint64_t>(); uint64_t* off = len + (S + 1);
            { ctx->launches++; k_shard_ins_len<<<(S + 1 + 255) / 256, 256, 0, st>>>(sorted, ctx->d_samp_idx.as<uint32_t>(), s_lo, S, tab, len); }
            size_t tmp = 0;
            cub::DeviceScan::ExclusiveSum(nullptr, tmp, len, off, (int)S + 1, st);
            SVIM_CUDA(ctx->d_sort_tmp.ensure(tmp));
            SVIM_CUDA(cub::DeviceScan::ExclusiveSum(ctx->d_sort_tmp.p, tmp, len, off, (int)S + 1, st));
            uint64_t total = 0;
            SVIM_CUDA(cudaMemcpyAsync(&total, off + S, 8, cudaMemcpyDeviceToHost, st));
            SVIM_CUDA(cudaStreamSynchronize(st));
            SVIM_CUDA(ctx->d_shard_ins.ensure((size_t)total + 16));
            if (total > 0) { ctx->launches++; k_shard_ins_copy<<<(uint32_t)(((uint64_t)S * 32 + 255) / 256), 256, 0, st>>>(ctx->d_csig_sorted.as<svim_csig>(), ctx->d_samp_idx.as<uint32_t>(), s_lo, S, tab, off,
                                                                                                                  ctx->d_shard_ins.as<uint8_t>()); }
            ins_blob = ctx->d_shard_ins.as<uint8_t>(); ins_blob_bytes = (int64_t)total;
        }
    }
    const uint32_t* d_small = ctx->d_plist.as<uint32_t>(); const uint32_t* d_large = d_small + P;
    const uint32_t* d_ins = d_large + P;
    uint32_t* d_misc = ctx->d_part_stats.as<uint32_t>();
    // ---- INS: edit distances ---------------------------------------------------------------------------------
    const int32_t* d_pair_ed = nullptr;
    if (n_ins > 0) {
        if (pair_total >= 0xffffffffull) { ctx->set_error(SVIMGPU_ERR_LIMIT, "too many insertion pairs"); return SVIMGPU_ERR_LIMIT; }
        SVIM_CUDA(ctx->d_pair_ed.ensure((size_t)pair_total * 4 + 4));
        const int64_t maxlen = ((ctx->cluster_max_ins_len + (int64_t)ceil(fabs(2.0 * cp.cluster_max_distance * cp.pos_norm)) + 64) + 15) & ~15ll;
        GenomeView gv{ctx->d_genome.as<uint8_t>(), ctx->d_genome_off.as<int64_t>(), ctx->genome_contigs, ctx->cluster_rank_to_tid, ctx->cluster_n_ranks};
        if (!ctx->d_genome.p) { gv.n = 0; }
        SVIM_CUDA(ctx->d_myers_ctl.ensure(MYERS_CTL_N * 4));
        uint32_t* d_ctl = ctx->d_myers_ctl.as<uint32_t>();   // list counts, hand-over capacities, cursors: see MYERS_CTL_* (myers.cu)
        uint32_t h_ctl[64] = {0};
        const uint32_t pblocks = (uint32_t)(((uint64_t)n_ins * 32 + 127) / 128);
        {
            StageTimer t(ctx, T_PAIRS);
            SVIM_CUDA(cudaMemsetAsync(d_ctl, 0, MYERS_CTL_N * 4, st));
            { ctx->launches++; k_ins_pairs<<<pblocks, 128, 0, st>>>(sorted, ins_blob, gv, ctx->d_samp_off.as<uint32_t>(), ctx->d_samp_idx.as<uint32_t>(), d_ins,
                                                n_ins, ctx->d_pair_off.as<uint64_t>(), cp, 0, nullptr, nullptr, d_ctl, d_ctl + MYERS_CTL_CAP,
                                                ctx->myers_band_num, ctx->myers_band_add, ctx->myers_tpp, d_misc + 9); }
            SVIM_CUDA(cudaMemcpyAsync(h_ctl, d_ctl, 64 * 4, cudaMemcpyDeviceToHost, st));
            SVIM_CUDA(cudaStreamSynchronize(st));
        }
        MyersPlan pl; memset(&pl, 0, sizeof(pl));
        uint32_t n_work = 0, n_banded = 0;
        uint32_t n_tpp = 0;
        for (int q = 0; q < MYERS_LISTS; ++q) { pl.cnt[q] = h_ctl[q]; pl.off[q] = n_work; n_work += h_ctl[q]; if (q >= MYERS_BINS) n_banded += h_ctl[q]; if (q >= 2 * MYERS_BINS) n_tpp += h_ctl[q]; }
        for (int bb = 0; bb < MYERS_BINS; ++bb) pl.retry_cap[bb] = h_ctl[MYERS_CTL_CAP + bb];
        cs.myers_pairs = n_work;
        if (n_work > 0 && !ctx->d_genome.p) { ctx->set_error(SVIMGPU_ERR_STATE, "insertion clustering needs svimgpu_set_genome"); return SVIMGPU_ERR_STATE; }
        if (n_work > 0) {
            SVIM_CUDA(ctx->d_pairs.ensure((size_t)n_work * 3 * sizeof(MyersWork) + 16));   // sorted list, fallback list, unsorted list (later: hand-over list)
            SVIM_CUDA(ctx->d_keys[0].ensure((size_t)n_work * 8)); SVIM_CUDA(ctx->d_keys[1].ensure((size_t)n_work * 8));
            MyersWork* d_unsorted = ctx->d_pairs.as<MyersWork>() + 2 * (size_t)n_work;
            MyersWork* d_work = ctx->d_pairs.as<MyersWork>();
            {
                StageTimer t(ctx, T_PAIRS);
                SVIM_CUDA(cudaMemcpyAsync(d_ctl, pl.off, MYERS_LISTS * 4, cudaMemcpyHostToDevice, st));
                { ctx->launches++; k_ins_pairs<<<pblocks, 128, 0, st>>>(sorted, ins_blob, gv, ctx->d_samp_off.as<uint32_t>(), ctx->d_samp_idx.as<uint32_t>(), d_ins,
                                                    n_ins, ctx->d_pair_off.as<uint64_t>(), cp, 1, d_unsorted, ctx->d_keys[0].as<uint64_t>(), d_ctl, d_ctl + MYERS_CTL_CAP,
                                                    ctx->myers_band_num, ctx->myers_band_add, ctx->myers_tpp, d_misc + 9); }
                // longest-processing-time-first inside every list: one radix sort on (list, ~cost)
                size_t tmp = 0;
                cub::DeviceRadixSort::SortPairs(nullptr, tmp, ctx->d_keys[0].as<uint64_t>(), ctx->d_keys[1].as<uint64_t>(), d_unsorted, d_work, (int)n_work, 0, 38, st);
                SVIM_CUDA(ctx->d_sort_tmp.ensure(tmp));
                SVIM_CUDA(cub::DeviceRadixSort::SortPairs(ctx->d_sort_tmp.p, tmp, ctx->d_keys[0].as<uint64_t>(), ctx->d_keys[1].as<uint64_t>(), d_unsorted, d_work,
                                                          (int)n_work, 0, 38, st));
            }
            StageTimer t(ctx, T_MYERS);
            int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
            MyersArgs ma; memset(&ma, 0, sizeof(ma));
            // symbol-code image of the INS blob for the thread-per-pair kernels (the genome's is built by svimgpu_set_genome)
            SVIM_CUDA(ctx->d_ins_codes.ensure((size_t)ins_blob_bytes + 16));
            if (ins_blob_bytes > 0) { ctx->launches++; k_tpp_encode<<<(unsigned)(((size_t)ins_blob_bytes + 16 * 256 - 1) / (16 * 256)), 256, 0, st>>>(ins_blob, ins_blob_bytes, ctx->d_ins_codes.as<uint8_t>()); }
            ma.genome_codes = ctx->d_genome_codes.as<uint8_t>(); ma.ins_codes = ctx->d_ins_codes.as<uint8_t>(); ma.ins_base = ins_blob;
            ma.sig = sorted; ma.ins_blob = ins_blob; ma.g = gv; ma.ed_out = ctx->d_pair_ed.as<int32_t>(); ma.maxlen = maxlen;
            ma.fallback = d_work + n_work; ma.cells = (unsigned long long*)(d_misc + 12); ma.err = d_misc + 9;
            ma.band_num = ctx->myers_band_num; ma.band_add = ctx->myers_band_add;
            StringPairs none{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
            // the unsorted list is dead after the sort: its room takes the banded pass's hand-overs
            SVIM_CUDA(myers_run_plan<false>(ctx, pl, ma, none, d_work, nullptr, d_unsorted, d_ctl, maxlen, sms));
            uint32_t h_retry[MYERS_CTL_FALLBACK - MYERS_CTL_RETRY];
            SVIM_CUDA(cudaMemcpyAsync(h_retry, d_ctl + MYERS_CTL_RETRY, sizeof(h_retry), cudaMemcpyDeviceToHost, st));
            SVIM_CUDA(cudaStreamSynchronize(st));
            uint32_t n_retry = 0; for (int bb = 0; bb < MYERS_BINS; ++bb) n_retry += h_retry[bb];
            cs.myers_banded_pairs = n_banded; cs.myers_retry_pairs = n_retry; cs.myers_tpp_pairs = n_tpp;
            uint64_t bc, tc; memcpy(&bc, h_retry + (MYERS_CTL_BANDCELLS - MYERS_CTL_RETRY), 8); memcpy(&tc, h_retry + (MYERS_CTL_TPPCELLS - MYERS_CTL_RETRY), 8);
            cs.myers_band_cells = (int64_t)(bc + tc);        // cells the first pass computed: wavefront bands + thread-per-pair windows
            cs.myers_tpp_cells = (int64_t)tc;
            uint64_t uc; memcpy(&uc, h_retry + (MYERS_CTL_UNBCELLS - MYERS_CTL_RETRY), 8); cs.myers_unbanded_cells = (int64_t)uc;
        }
        d_pair_ed = ctx->d_pair_ed.as<int32_t>();
    }
    // ---- linkage --------------------------------------------------------------------------------------------------
    SVIM_CUDA(ctx->d_labels.ensure((size_t)n_samp * 4 + 4)); SVIM_CUDA(ctx->d_part_ncl.ensure((size_t)(P + 1) * 4)); SVIM_CUDA(ctx->d_part_nkept.ensure((size_t)(P + 1) * 4));
    SVIM_CUDA(cudaMemsetAsync(ctx->d_part_ncl.p, 0, (size_t)(P + 1) * 4, st)); SVIM_CUDA(cudaMemsetAsync(ctx->d_part_nkept.p, 0, (size_t)(P + 1) * 4, st));
    {
        StageTimer t(ctx, T_LINKAGE);
        LinkArgs la{sorted, ctx->d_samp_off.as<uint32_t>(), ctx->d_samp_idx.as<uint32_t>(), nullptr, 0, ctx->d_pair_off.as<uint64_t>(), d_pair_ed, cp,
                    ctx->d_labels.as<int32_t>(), ctx->d_part_ncl.as<uint32_t>(), ctx->d_part_nkept.as<uint32_t>(), d_misc + 9, ctx->d_hdr.as<uint32_t>() + HDR_DUP_T};
        if (n_small) {
            la.plist = d_small; la.n_list = n_small;
            { ctx->launches++; k_linkage<<<la.n_list, 32, part_smem_bytes(max_m_small), st>>>(la, max_m_small); }
        }
        if (n_big) {
            size_t sm = part_smem_bytes(100);
            SVIM_CUDA(cudaFuncSetAttribute(k_linkage, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
            la.plist = d_large; la.n_list = n_big;
            { ctx->launches++; k_linkage<<<la.n_list, 32, sm, st>>>(la, 100); }
        }
        SVIM_CUDA(cudaGetLastError());
    }
    // ---- consolidate ----------------------------------------------------------------------------------------------
    {
        StageTimer t(ctx, T_CONSOLIDATE);
        SVIM_CUDA(ctx->d_cl_off.ensure((size_t)(P + 1) * 4)); SVIM_CUDA(ctx->d_mem_off.ensure((size_t)(P + 1) * 4));
        size_t tmp = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tmp, ctx->d_part_ncl.as<uint32_t>(), ctx->d_cl_off.as<uint32_t>(), (int)P + 1, st);
        SVIM_CUDA(ctx->d_sort_tmp.ensure(tmp));
        SVIM_CUDA(cub::DeviceScan::ExclusiveSum(ctx->d_sort_tmp.p, tmp, ctx->d_part_ncl.as<uint32_t>(), ctx->d_cl_off.as<uint32_t>(), (int)P + 1, st));
        SVIM_CUDA(cub::DeviceScan::ExclusiveSum(ctx->d_sort_tmp.p, tmp, ctx->d_part_nkept.as<uint32_t>(), ctx->d_mem_off.as<uint32_t>(), (int)P + 1, st));
        uint32_t tot[2];
        SVIM_CUDA(cudaMemcpyAsync(&tot[0], ctx->d_cl_off.as<uint32_t>() + P, 4, cudaMemcpyDeviceToHost, st));
        SVIM_CUDA(cudaMemcpyAsync(&tot[1], ctx->d_mem_off.as<uint32_t>() + P, 4, cudaMemcpyDeviceToHost, st));
        SVIM_CUDA(cudaStreamSynchronize(st));
        n_clusters = tot[0]; n_members = tot[1];
        SVIM_CUDA(ctx->d_clusters.ensure((size_t)(n_clusters + 1) * sizeof(svim_cluster)));
        SVIM_CUDA(ctx->d_clusters_sorted.ensure((size_t)(n_clusters + 1) * sizeof(svim_cluster)));
        SVIM_CUDA(ctx->d_members.ensure((size_t)(n_members + 1) * 4));
        SVIM_CUDA(cudaMemsetAsync(ctx->d_clusters.p, 0, (size_t)(n_clusters + 1) * sizeof(svim_cluster), st));
        if (n_clusters > 0) {
            { ctx->launches++; k_members<<<(uint32_t)(((uint64_t)P * 32 + 127) / 128), 128, 0, st>>>(ctx->d_samp_off.as<uint32_t>(), ctx->d_samp_idx.as<uint32_t>(),
                ctx->d_order.as<uint32_t>(), ctx->d_labels.as<int32_t>(), ctx->d_part_ncl.as<uint32_t>(), ctx->d_cl_off.as<uint32_t>(),
                ctx->d_mem_off.as<uint32_t>(), P, ctx->d_clusters.as<svim_cluster>(), ctx->d_members.as<uint32_t>()); }
            { ctx->launches++; k_consolidate<<<(n_clusters + 63) / 64, 64, 0, st>>>(ctx->d_csig.as<svim_csig>(), ctx->d_members.as<uint32_t>(), n_clusters,
                                                               ctx->d_clusters.as<svim_cluster>(), d_misc + 9); }
        }
        SVIM_CUDA(cudaGetLastError());
    }
    if (shard_n > 1) {          // data errors of this shard (zero span, mixed BND directions, haplotype bound) count as a failure of the rank
        uint32_t flag = 0;
        SVIM_CUDA(cudaMemcpyAsync(&flag, ctx->d_part_stats.as<uint32_t>() + 9, 4, cudaMemcpyDeviceToHost, st));
        SVIM_CUDA(cudaStreamSynchronize(st));
        if (flag) { ctx->set_error(SVIMGPU_ERR_DATA, "cluster: data error %u in this rank's partitions", flag); return SVIMGPU_ERR_DATA; }
    }
    return 0;
    };
    const int rc_local = local();
    if (stop_after_partition) { if (stats) *stats = cs; return rc_local; }
    if (shard_n > 1) { int rc = cluster_exchange(ctx, &n_clusters, &n_members, rc_local); if (rc) return rc; }
    else if (rc_local) return rc_local;
    uint32_t* d_misc = ctx->d_part_stats.as<uint32_t>();
    // ---- final order + D2H -------------------------------------------------------------------------------------------
    SVIM_CUDA(ctx->h_clusters.ensure((size_t)(n_clusters + 1) * sizeof(svim_cluster))); SVIM_CUDA(ctx->h_members.ensure((size_t)(n_members + 1) * 4));
    ctx->n_clusters_host = n_clusters; ctx->n_members_host = n_members;
    // the member array is final here: its D2H runs on a side stream under the ordering of the cluster records
    cudaStream_t side = ctx->aux_stream[0];
    if (n_members) {
        SVIM_CUDA(cudaEventRecord(ctx->aux_ev[0], st));
        SVIM_CUDA(cudaStreamWaitEvent(side, ctx->aux_ev[0], 0));
        SVIM_CUDA(cudaMemcpyAsync(ctx->h_members.p, ctx->d_members.p, (size_t)n_members * 4, cudaMemcpyDeviceToHost, side));
    }
    if (n_clusters > 0) {
        StageTimer t(ctx, T_ORDER);
        const uint32_t cb = (n_clusters + 255) / 256;
        for (int k = 0; k < 2; ++k) { SVIM_CUDA(ctx->d_ckeys[k].ensure((size_t)n_clusters * 8)); SVIM_CUDA(ctx->d_cvals[k].ensure((size_t)n_clusters * 4)); }
        SVIM_CUDA(ctx->d_keys[0].ensure((size_t)n_clusters * 8));
        uint64_t* k2 = ctx->d_keys[0].as<uint64_t>();
        { ctx->launches++; k_final_keys<<<cb, 256, 0, st>>>(ctx->d_clusters.as<svim_cluster>(), ctx->d_csig.as<svim_csig>(), ctx->d_members.as<uint32_t>(), n_clusters,
                                         ctx->d_ckeys[0].as<uint64_t>(), k2, ctx->d_cvals[0].as<uint32_t>(), ctx->d_hdr.as<uint32_t>() + HDR_NCL_T); }
        uint64_t* k_out; uint32_t* v_out;
        int rc = sort_pairs_u64(ctx, ctx->d_ckeys, ctx->d_cvals, n_clusters, 64, &k_out, &v_out); if (rc) return rc;
        { ctx->launches++; k_gather<uint64_t><<<cb, 256, 0, st>>>(k2, v_out, n_clusters, ctx->d_ckeys[1].as<uint64_t>()); }
        std::swap(ctx->d_ckeys[0], ctx->d_ckeys[1]);
        rc = sort_pairs_u64(ctx, ctx->d_ckeys, ctx->d_cvals, n_clusters, 64, &k_out, &v_out); if (rc) return rc;
        { ctx->launches++; k_gather<svim_cluster><<<cb, 256, 0, st>>>(ctx->d_clusters.as<svim_cluster>(), v_out, n_clusters, ctx->d_clusters_sorted.as<svim_cluster>()); }
        SVIM_CUDA(cudaGetLastError());
    }
    uint32_t* hh = ctx->h_hdr.as<uint32_t>();          // [0, HDR_WORDS) plan header incl. per-type counters, then 16 words of d_misc
    {
        StageTimer t(ctx, T_CLUSTER_D2H);              // page-locked destinations: the copies run at PCIe rate and do not stage through the driver
        if (n_clusters) SVIM_CUDA(cudaMemcpyAsync(ctx->h_clusters.p, ctx->d_clusters_sorted.p, (size_t)n_clusters * sizeof(svim_cluster), cudaMemcpyDeviceToHost, st));
        SVIM_CUDA(cudaMemcpyAsync(hh, ctx->d_hdr.p, HDR_WORDS * 4, cudaMemcpyDeviceToHost, st));
        SVIM_CUDA(cudaMemcpyAsync(hh + HDR_WORDS, d_misc, 16 * 4, cudaMemcpyDeviceToHost, st));
    }
    SVIM_CUDA(cudaStreamSynchronize(st));
    if (n_members) SVIM_CUDA(cudaStreamSynchronize(side));
    const uint32_t* h = hh + HDR_WORDS;
    if (h[9]) {
        const char* why = h[9] == 2 ? "haplotype longer than the scratch bound" : h[9] == 3 ? "BND cluster with mixed directions (assertion in consolidate_clusters_bilocal)"
                                    : "division by zero span (ZeroDivisionError in the reference) or contig missing from the genome";
        ctx->set_error(SVIMGPU_ERR_DATA, "cluster: %s", why);
        return SVIMGPU_ERR_DATA;
    }
    unsigned long long cells; memcpy(&cells, h + 12, 8);
    cs.myers_cells = (int64_t)cells;
    for (int ty = 0; ty < 6; ++ty) { cs.n_clusters[ty] = hh[HDR_NCL_T + ty]; cs.duplicate_signatures[ty] = hh[HDR_DUP_T + ty]; }   // duplicates: this rank's partitions
    cs.n_clusters_total = n_clusters; cs.n_members = n_members;
    if (stats) *stats = cs;
    return 0;
}
