// scl_verify.cuh -- STRICT mode, lists 17..32: check a recorded close decision in double instead of decoding the whole
// codeword again.
//
// The first pass (scl_fast.cuh) records a keep/drop decision whose margin is below tau when exactly one kept fork K
// and one dropped fork D lie near the cut (close_decision): the leaf phi, the two forks (lane, bit) and the decided
// bits of their parent paths below phi. Every decision before phi had a safe margin or was itself recorded, so these
// prefixes are the reference's own paths, and what the reference compares at phi (PolarCode.cpp:497-553) is
//     metric(fork) = sum over the leaves below phi of log(1 + exp(-+LLR_leaf)) + log(1 + exp(-+LLR_phi))
// along each parent. With all decisions known the leaf LLRs have no serial dependency: the partial sums of every subtree
// follow from the bits (the encoder recursion, :457-473 run for every node), and each tree layer is one parallel step
// over all N positions (:422-455 for every node of the layer). One block does that for both parents in double -- 2 x
// N log N node updates, log N deep, against the N-deep chain of a full decode -- and compares the two metrics under the
// reference's order (metric, then fork index). Only if the reference would have kept D instead (or the two metrics
// agree to 1e-9, an exact cancellation) is the codeword appended to the flag list for the full second pass.
//
// Included by polar_b200.cu after scl_exact.cuh (double-precision f / softplus with short dependency chains).
#pragma once

namespace verify {

constexpr int NT = 256;

struct Args {
    const float* llr;            // [B][N] channel LLRs of the batch the records refer to
    const uint32_t* vrec;        // records: 4 + 2 * NW words each (fastcommon::Args::vrec)
    const int* vcount;           // number of records written (may exceed vcap: the overflow was flagged by the first pass)
    int vcap;
    const uint32_t* frozen_words;
    int* flag_list;
    int* flag_count;
    int* cw_state;               // [B]: 1 once a codeword is on the flag list (no duplicates)
    double* gap_out;             // [vcap] or null: metric(D) - metric(K) in double per record (cross-check against rec[3])
    int n;
};

// u-hat of a block of M bits starting at bit `off` of x[] (packed): the packed polar transform the first pass uses for
// its own output (strides M/2 .. 1, a[i] ^= a[i + s] for i with (i & s) == 0), in place. One warp.
__device__ __forceinline__ void block_transform(uint32_t* x, int off, int M, int lane) {
    if (M >= 32) {
        const int w0 = off >> 5, mw = M >> 5;
        for (int s = mw >> 1; s >= 1; s >>= 1) {
            for (int i = lane; i < mw; i += 32)
                if ((i & s) == 0) x[w0 + i] ^= x[w0 + i + s];
            __syncwarp();
        }
        for (int i = lane; i < mw; i += 32) {
            uint32_t v = x[w0 + i];
            v ^= (v >> 16) & 0x0000FFFFu;
            v ^= (v >> 8) & 0x00FF00FFu;
            v ^= (v >> 4) & 0x0F0F0F0Fu;
            v ^= (v >> 2) & 0x33333333u;
            v ^= (v >> 1) & 0x55555555u;
            x[w0 + i] = v;
        }
        __syncwarp();
    } else if (lane == 0) {
        const uint32_t mask = (M == 32) ? 0xFFFFFFFFu : ((1u << M) - 1u);
        uint32_t f = (x[off >> 5] >> (off & 31)) & mask;
        for (int s = M >> 1; s >= 1; s >>= 1) {
            uint32_t sel = 0;                                    // bits i of the field with (i & s) == 0
            for (int i = 0; i < M; ++i) if ((i & s) == 0) sel |= 1u << i;
            f ^= (f >> s) & sel;
        }
        x[off >> 5] = (x[off >> 5] & ~(mask << (off & 31))) | (f << (off & 31));
    }
}

__global__ void __launch_bounds__(NT) verify_kernel(const Args a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
    const int n = a.n, N = 1 << n, NW = N >> 5;
    // shared: two LLR layers (double), partial sums of every layer (bit packed, node-major), reduction scratch, tables
    double* La = reinterpret_cast<double*>(smem_raw);
    double* Lb = La + N;
    uint32_t* Bl = reinterpret_cast<uint32_t*>(Lb + N);           // [(n + 1)][NW]; layer n = the decided bits u
    double* red = reinterpret_cast<double*>(Bl + (size_t)(n + 1) * NW);   // [NT / 32 + 2]
    exact::Tables* tb = reinterpret_cast<exact::Tables*>(red + NT / 32 + 2);
    exact::build_tables(tb, tid, NT);
    const int nrec = min(*a.vcount, a.vcap);
    const int rec_words = 4 + 2 * NW;
    for (int r = blockIdx.x; r < nrec; r += gridDim.x) {
        const uint32_t* rec = a.vrec + (size_t)r * rec_words;
        const int cw = (int)rec[0], phi = (int)rec[1];
        const int lK = rec[2] & 0xFF, bK = (rec[2] >> 8) & 1, lD = (rec[2] >> 16) & 0xFF, bD = (rec[2] >> 24) & 1;
        const float* chan = a.llr + (size_t)cw * N;
        double metric[2];
#pragma unroll 1
        for (int which = 0; which < 2; ++which) {
            const int bit = which ? bD : bK;
            __syncthreads();
            // ---- the parent's decided bits below phi ----
            uint32_t* U = Bl + (size_t)n * NW;
            for (int i = tid; i < NW; i += NT) U[i] = 0u;
            __syncthreads();
            const uint32_t* src = rec + 4 + which * NW;
            for (int lam = 1; lam <= n; ++lam) {
                if (!((phi >> (n - lam)) & 1)) continue;
                const int M = N >> lam, off = (phi >> (n - lam + 1)) << (n - lam + 1);
                if (M >= 32) { for (int i = tid; i < (M >> 5); i += NT) U[(off >> 5) + i] = src[(off >> 5) + i]; }
                else if (tid == 0) {
                    const uint32_t mask = ((1u << M) - 1u) << (off & 31);
                    U[off >> 5] = (U[off >> 5] & ~mask) | (src[off >> 5] & mask);
                }
                __syncthreads();
            }
            // each completed block back to decided bits (its own polar transform); blocks go round the warps
            {
                int k = 0;
                for (int lam = 1; lam <= n - 5; ++lam) {           // blocks of whole words: round the warps
                    if (!((phi >> (n - lam)) & 1)) continue;
                    if ((k++ % (NT / 32)) == wib) block_transform(U, (phi >> (n - lam + 1)) << (n - lam + 1), N >> lam, lane);
                }
                if (wib == 0)                                      // the blocks inside phi's own word: one warp, in turn
                    for (int lam = (n - 4 > 1 ? n - 4 : 1); lam <= n; ++lam)
                        if ((phi >> (n - lam)) & 1) { block_transform(U, (phi >> (n - lam + 1)) << (n - lam + 1), N >> lam, lane); __syncwarp(); }
            }
            __syncthreads();
            // ---- partial sums of every node, bottom-up (PolarCode.cpp:457-473 for all nodes): layer lam - 1, node p,
            // entries (2b, 2b + 1) = (B[lam][2p][b] ^ B[lam][2p + 1][b], B[lam][2p + 1][b]); one output word per thread
            for (int lam = n; lam >= 2; --lam) {
                const int M = N >> lam;                           // entries per node at layer lam
                const uint32_t* in = Bl + (size_t)lam * NW;
                uint32_t* out = Bl + (size_t)(lam - 1) * NW;
                for (int wd = tid; wd < NW; wd += NT) {
                    uint32_t o = 0;
                    const int t0 = wd << 5;                       // first output index of this word (node-major)
                    for (int j = 0; j < 32; j += 2) {
                        const int t = t0 + j;                     // even output index 2b of node p
                        const int p = t / (2 * M), b = (t % (2 * M)) >> 1;
                        const int ia = (2 * p) * M + b, ib = (2 * p + 1) * M + b;
                        const uint32_t xa = (in[ia >> 5] >> (ia & 31)) & 1u, xb = (in[ib >> 5] >> (ib & 31)) & 1u;
                        o |= ((xa ^ xb) << j) | (xb << (j + 1));
                    }
                    out[wd] = o;
                }
                __syncthreads();
            }
            // ---- LLR layers top-down, every node of a layer at once (PolarCode.cpp:422-455) ----
            for (int i = tid; i < N; i += NT) La[i] = (double)chan[i];
            __syncthreads();
            double* prev = La;
            double* cur = Lb;
            for (int lam = 1; lam <= n; ++lam) {
                const int M = N >> lam;
                const uint32_t* Bs = Bl + (size_t)lam * NW;
                for (int t = tid; t < N; t += NT) {
                    const int p = t / M, b = t % M;
                    const double* x = prev + (size_t)(p >> 1) * (2 * M) + 2 * b;
                    double y;
                    if (p & 1) {
                        const int ib = (p - 1) * M + b;
                        const uint32_t u = (Bs[ib >> 5] >> (ib & 31)) & 1u;
                        y = x[1] + (u ? -x[0] : x[0]);
                    } else {
                        y = exact::f_fast(x[0], x[1], tb);
                    }
                    cur[t] = y;
                }
                __syncthreads();
                double* tmp = prev; prev = cur; cur = tmp;
            }
            // prev[leaf] = decision LLR of every leaf along this parent. Metric of the fork = sum below phi + the fork's own term
            double acc = 0.0;
            for (int leaf = tid; leaf <= phi; leaf += NT) {
                const bool frozen = (a.frozen_words[leaf >> 5] >> (leaf & 31)) & 1u;
                const uint32_t u = (leaf == phi) ? (uint32_t)bit : ((U[leaf >> 5] >> (leaf & 31)) & 1u);
                acc += exact::softplus_fast((!frozen && u) ? prev[leaf] : -prev[leaf], tb);
            }
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(FULL_MASK, acc, o);
            if (lane == 0) red[wib] = acc;
            __syncthreads();
            if (tid == 0) {
                double t = 0.0;
                for (int i = 0; i < NT / 32; ++i) t += red[i];
                red[NT / 32 + which] = t;
            }
            __syncthreads();
            metric[which] = red[NT / 32 + which];
        }
        if (tid == 0) {
            const double mK = metric[0], mD = metric[1];
            const int iK = 2 * lK + bK, iD = 2 * lD + bD;
            if (a.gap_out != nullptr) a.gap_out[r] = mD - mK;
            const bool kept_is_right = (mK < mD || (mK == mD && iK < iD)) && fabs(mK - mD) >= exact::kTieMargin;
            if (!kept_is_right && atomicExch(a.cw_state + cw, 1) == 0) a.flag_list[atomicAdd(a.flag_count, 1)] = cw;
        }
    }
}

}  // namespace verify
