// ssw.cu -- sm_100a kernels for batched ssw_init + ssw_align on reference windows
// (ssw.c:742 ssw_init, :771 ssw_align, :371 sw_sse2_word, :549 banded_sw), as salt's
// mate rescue calls them (alnpe.c:261 snpaln_sw_snpaware, :330 snpaln_sw).
//
// Pipeline per batch (all on the caller's stream, no host round trip):
//   sw_prep<fwd>   window symbols + read-code selectors per task
//   sw_dp<fwd>     score1 / ref_end1 / read_end1 / maxColumn -> score2 / ref_end2
//   sw_prep<rev>   reversed read prefix and reversed window (ssw.c:827-832)
//   sw_dp<rev>     ref_begin1 / read_begin1, stops at the column whose maximum equals score1
//   sw_banded      banded_sw + traceback -> cigar (ssw.c:549-727); overflow pass for wide bands
#if !defined(SALT_EMUL)
#include <cuda_runtime.h>
#endif

#include "common.cuh"
#include "sw_core.cuh"
#include "kernels.h"

namespace salt {

// fwd[] record per task (int32 x 8)
enum { F_SCORE1 = 0, F_REF_END1, F_READ_END1, F_SCORE2, F_REF_END2, F_REF_BEGIN1, F_READ_BEGIN1, F_FLAGS };
enum { FL_VALID = 1, FL_DO_REV = 2, FL_DO_CIGAR = 4 };

struct SwDev {
    DevCtx c;
    const salt_win_t *wins;
    size_t n_tasks;            // real tasks
    size_t n_pairs;            // ceil(n_tasks / 2)
    int G, S;                  // rows = G*S
    int CW;                    // window words (8 columns each) per pair
    int item_shift;            // log2 of the prep kernels' padded items per task
    int MC;                    // maxcol stride (columns) per pair
    uint2 *win2;               // [pair][CW]  (.x task 0, .y task 1)
    uint32_t *maxcol2;         // [pair][MC]  packed column maxima
    int32_t *fwd;              // [task][8]
    SswParams prm;
};

__device__ __forceinline__ int sw_ref_symbol(const DevCtx &c, int use_pac, uint32_t p)
{
    if (use_pac) return (c.pac[p >> 2] >> ((~p & 3u) << 1)) & 3;
    return (c.mixref[p >> 3] >> (4u * (p & 7u))) & 15u;
}

// eight consecutive pac symbols starting at position p, one per nibble (lowest nibble first)
__device__ __forceinline__ uint32_t sw_pac_group(const uint8_t *__restrict__ pac, uint32_t p)
{
    const uint8_t *__restrict__ b = pac + (p >> 2);
    const uint32_t x = ((uint32_t)b[0] << 16) | ((uint32_t)b[1] << 8) | (uint32_t)b[2];      // 12 symbols, first in the top bits
    uint32_t r = 0;
#pragma unroll
    for (int m = 0; m < 8; ++m) r |= ((x >> (22 - 2 * (int)((p & 3u) + (uint32_t)m))) & 3u) << (4 * m);
    return r;
}

// read code 0..4 of packed read rs at index i (one-hot nibble -> code)
__device__ __forceinline__ int sw_read_code(const DevCtx &c, uint32_t rs, int i)
{
    const uint64_t w = c.rd4[(size_t)rs * c.W64 + (i >> 4)];
    const unsigned nib = (unsigned)(w >> (4 * (i & 15))) & 15u;
    return nib == 15u ? SW_CODE_N : (31 - __clz(nib));
}

// --------------------------------------------------------------------------------------
// prep: one thread per window word (eight reference symbols of a task); the read-code selectors are made by the DP kernel.
// --------------------------------------------------------------------------------------
template <bool REV>
__global__ void __launch_bounds__(256)
sw_prep_kernel(SwDev d)
{
    const int per_task = d.CW;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t t = idx >> d.item_shift;              // a task's items are padded to a power of two: no division
    const int item = (int)(idx & (((size_t)1 << d.item_shift) - 1));
    if (t >= d.n_pairs * 2 || item >= per_task) return;
    const bool real = t < d.n_tasks;
    uint32_t rs = 0, start = 0;
    int L = 0, cols = 0, read_end = 0, ref_end = 0;
    int32_t *f = d.fwd + t * 8;
    if (real) {
        const salt_win_t w = d.wins[t];
        const uint32_t rid = w.rs >> 1;
        const int64_t lim = d.prm.use_pac ? d.c.l_pac : (int64_t)d.c.l;
        const bool ok = rid < d.c.n_reads && w.start <= w.end && (int64_t)w.end <= lim &&      // end == l is what alnpe.c:213-252 clamps to: position l reads the zero pad
                        (!d.prm.use_pac || d.c.pac != nullptr) &&
                        (int64_t)(w.end - w.start + 1) <= (int64_t)d.MC;
        if (!REV) {
            if (item == 0) {
                f[F_FLAGS] = ok ? FL_VALID : 0;
                f[F_SCORE1] = 0; f[F_REF_END1] = 0; f[F_READ_END1] = 0; f[F_SCORE2] = 0;
                f[F_REF_END2] = 0;
                f[F_REF_BEGIN1] = -1; f[F_READ_BEGIN1] = -1;
            }
            if (ok) { rs = w.rs; start = w.start; L = d.c.rd_len[rid]; cols = (int)(w.end - w.start + 1); }
        } else {
            if (ok && (f[F_FLAGS] & FL_DO_REV)) {
                rs = w.rs; start = w.start; read_end = f[F_READ_END1]; ref_end = f[F_REF_END1];
                L = read_end + 1; cols = ref_end + 1;
            }
        }
    }
    {                                                   // window word: columns 8*item .. 8*item+7
        uint32_t word = 0;
        const int col0 = item * 8;
        if (!d.prm.use_pac && col0 < cols) {
            // eight mask nibbles with one funnel shift of two aligned reference words (the reference is
            // padded on both sides, engine.cu); the reverse pass wants them in descending order
            const int nv = cols - col0 < 8 ? cols - col0 : 8;
            const int64_t p0 = REV ? (int64_t)start + ref_end - col0 - 7 : (int64_t)start + col0;   // lowest position of the group
            const uint32_t *__restrict__ mw = d.c.mixref + (p0 >> 3);
            uint32_t x = __funnelshift_r(mw[0], mw[1], 4 * (int)(p0 & 7));
            if (REV) {
                x = __byte_perm(x, 0u, 0x0123);                                   // byte order, then the nibbles in each byte
                x = ((x & 0x0F0F0F0Fu) << 4) | ((x >> 4) & 0x0F0F0F0Fu);
            }
            word = nv < 8 ? (x & ((1u << (4 * nv)) - 1u)) : x;
        } else {
            for (int b = 0; b < 8; ++b) {
                const int col = col0 + b;
                if (col < cols) {
                    const uint32_t p = REV ? start + (uint32_t)(ref_end - col) : start + (uint32_t)col;
                    word |= (uint32_t)sw_ref_symbol(d.c, d.prm.use_pac, p) << (4 * b);
                }
            }
        }
        uint32_t *dst = reinterpret_cast<uint32_t *>(d.win2 + (t >> 1) * d.CW + item);
        dst[t & 1] = word;
    }
}

// --------------------------------------------------------------------------------------
// dp: systolic strip DP, see sw_core.cuh.  CTA = 128 threads = 128/G task pairs.
// --------------------------------------------------------------------------------------
__device__ __forceinline__ int half_of(uint32_t x, int h) { return h ? s16hi(x) : s16lo(x); }

// keep a loop-invariant value in its register: without this ptxas re-derives it from the kernel
// parameters inside the column loop to save a register, which costs more issue slots than it saves
#if defined(__CUDA_ARCH__)
#define SALT_PIN32(x) asm volatile("" : "+r"(x))
#define SALT_PIN64(x) asm volatile("" : "+l"(x))
#else
#define SALT_PIN32(x) (void)(x)
#define SALT_PIN64(x) (void)(x)
#endif

template <int G, int S, bool REV>
__global__ void __launch_bounds__(128, (G == 4 && S == 26) ? 5 : 1)
sw_dp_kernel(SwDev d)
{
    constexpr int NQ = (S + 3) / 4;
    constexpr int SP = (S + 3) / 4 * 4;                  // snapshot row stride: whole uint4s
    constexpr unsigned FULL = 0xffffffffu;
    __shared__ uint2 s_tab[17];
    SALT_DYN_SMEM(uint32_t, s_snap);                     // [thread][2][SP]
    if (threadIdx.x < 17) {
        const uint32_t *tb = reinterpret_cast<const uint32_t *>(d.prm.table);
        s_tab[threadIdx.x] = make_uint2(tb[2 * threadIdx.x], tb[2 * threadIdx.x + 1]);
    }
    __syncthreads();
    constexpr int GPC = 128 / G;
    const int gl = threadIdx.x / G, j = threadIdx.x % G;
    const size_t pair_raw = (size_t)blockIdx.x * GPC + gl;
    const bool live = pair_raw < d.n_pairs;              // dead groups run along on empty tasks (warp-uniform flow)
    const size_t pair = live ? pair_raw : 0;
    const int gshift = (threadIdx.x & 31) / G * G;
    uint32_t *snap = s_snap + (size_t)threadIdx.x * 2 * SP;

    int rows[2], cols[2], off[2], aux_read_end[2], aux_ref_end[2];
    uint32_t term = 0, task_rs[2] = {0u, 0u};
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const size_t t = pair * 2 + h;
        rows[h] = 0; cols[h] = 0; aux_read_end[h] = 0; aux_ref_end[h] = 0;
        if (live && t < d.n_tasks) {
            const int32_t *f = d.fwd + t * 8;
            const int fl = f[F_FLAGS];
            if (!REV) {
                if (fl & FL_VALID) {
                    const salt_win_t w = d.wins[t];
                    task_rs[h] = w.rs;
                    rows[h] = d.c.rd_len[w.rs >> 1];
                    cols[h] = (int)(w.end - w.start + 1);
                }
            } else if (fl & FL_DO_REV) {
                task_rs[h] = d.wins[t].rs;
                aux_read_end[h] = f[F_READ_END1]; aux_ref_end[h] = f[F_REF_END1];
                rows[h] = aux_read_end[h] + 1; cols[h] = aux_ref_end[h] + 1;
                term |= (uint32_t)(uint16_t)(f[F_SCORE1] + SW_BIAS) << (16 * h);      // compared in the biased domain
            }
        }
        off[h] = G * S - 8 * ((rows[h] + 7) / 8);
    }
    int maxcols = cols[0] > cols[1] ? cols[0] : cols[1];

    SwStrip<S> st;
    st.clear();
    // read-code selectors of this thread's S rows (four rows per word): rows above the read score -128 against everything,
    // the read's own rows carry its codes (the reverse pass walks the read backwards from read_end1, ssw.c:827-832), the
    // padded rows [L, 8*ceil(L/8)) score 0
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int Lh = rows[h];
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            uint32_t sel = 0;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int i = 4 * q + r;
                int code = SW_CODE_TOP;
                if (i < S) {
                    const int row = j * S + i - off[h];
                    if (row >= 0) code = row < Lh ? sw_read_code(d.c, task_rs[h], REV ? aux_read_end[h] - row : row) : SW_CODE_TAIL;
                }
                sel |= (uint32_t)code << (4 * r);
            }
            if (h == 0) st.sel0[q] = sel; else st.sel1[q] = sel;
        }
    }
    uint32_t negO = s16x2(-d.prm.gapO, -d.prm.gapO);
    uint32_t negE = 0u - ((uint32_t)d.prm.gapE | ((uint32_t)d.prm.gapE << 16));   // one 32-bit subtraction for both halves
    SALT_PIN32(negO); SALT_PIN32(negE);
    const uint2 *__restrict__ win = d.win2 + pair * d.CW;
    uint32_t *__restrict__ mcol = d.maxcol2 + pair * (size_t)d.MC;
    SALT_PIN64(win); SALT_PIN64(mcol);

    uint32_t Hrecv = SW_BIAS2, Hrecv2 = SW_BIAS2, Frecv = SW_BIAS2, Crecv = SW_BIAS2, best = SW_BIAS2;
    int bcol[2] = {0, 0};
    uint2 wreg = make_uint2(0, 0);
    int done = 0;                                        // REV, lane G-1: bit h = half h reached score1
    if (REV) done = (cols[0] == 0 ? 1 : 0) | (cols[1] == 0 ? 2 : 0);
    const int nsteps = __reduce_max_sync(FULL, maxcols + G - 1);    // warp-uniform: every shuffle runs on the full mask
    for (int t = 0; t < nsteps; ++t) {
        const int c = t - j;
        uint32_t Hbot = SW_BIAS2, Fout = SW_BIAS2, cm = SW_BIAS2;
        if (c >= 0 && c < maxcols) {
            if ((c & 7) == 0) wreg = win[c >> 3];
            const int sh = 4 * (c & 7);
            int sym0 = (wreg.x >> sh) & 15, sym1 = (wreg.y >> sh) & 15;
            if (c >= cols[0]) sym0 = SW_SYM_PADCOL;
            if (c >= cols[1]) sym1 = SW_SYM_PADCOL;
            const uint2 ta = s_tab[sym0], tb = s_tab[sym1];
            uint32_t F = j == 0 ? SW_BIAS2 : Frecv;
            const uint32_t diag = j == 0 ? SW_BIAS2 : Hrecv2;
            const uint32_t sm = st.column(ta.x, ta.y, tb.x, tb.y, diag, F, negO, negE);
            Hbot = st.H[S - 1]; Fout = F;
            cm = vmax2(j == 0 ? SW_BIAS2 : Crecv, sm);
            const uint32_t nb = vmax2(best, sm);
            if (nb != best) {                            // this strip's maximum rose: remember where
                const uint32_t ch = nb ^ best;
                if (ch & 0xffffu) {
                    bcol[0] = c;
#pragma unroll
                    for (int i = 0; i < SP; i += 4)
                        *reinterpret_cast<uint4 *>(snap + i) = make_uint4(st.H[i], i + 1 < S ? st.H[i + 1 < S ? i + 1 : 0] : 0u,
                                                                          i + 2 < S ? st.H[i + 2 < S ? i + 2 : 0] : 0u,
                                                                          i + 3 < S ? st.H[i + 3 < S ? i + 3 : 0] : 0u);
                }
                if (ch >> 16) {
                    bcol[1] = c;
#pragma unroll
                    for (int i = 0; i < SP; i += 4)
                        *reinterpret_cast<uint4 *>(snap + SP + i) = make_uint4(st.H[i], i + 1 < S ? st.H[i + 1 < S ? i + 1 : 0] : 0u,
                                                                               i + 2 < S ? st.H[i + 2 < S ? i + 2 : 0] : 0u,
                                                                               i + 3 < S ? st.H[i + 3 < S ? i + 3 : 0] : 0u);
                }
                best = nb;
            }
            if (j == G - 1) {
                if (!REV) mcol[c] = cm - SW_BIAS2;       // no borrow: both halves >= the bias
                else {
                    if (c < cols[0] && (cm & 0xffffu) == (term & 0xffffu)) done |= 1;
                    if (c < cols[1] && (cm >> 16) == (term >> 16)) done |= 2;
                }
            }
        }
        Hrecv2 = Hrecv;
        Hrecv = __shfl_up_sync(FULL, Hbot, 1, G);
        Frecv = __shfl_up_sync(FULL, Fout, 1, G);
        Crecv = __shfl_up_sync(FULL, cm, 1, G);
        if (REV) {
            // a task pair that reached score1 on both halves stops computing (ssw.c:500); the warp leaves
            // the loop when all of its pairs have
            const bool gdone = __shfl_sync(FULL, done, G - 1, G) == 3;
            if (gdone) maxcols = 0;
            if (__all_sync(FULL, gdone)) break;
        }
    }
    __syncwarp();
    // the epilogue branches per task pair (mask length, flags): group masks from here on
    const unsigned gmask = (unsigned)(((1ull << G) - 1ull) << gshift);

    // ---- combine the strips: global maximum, first column reaching it, smallest row there
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const size_t t = pair * 2 + h;
        const int bh = half_of(best, h) - SW_BIAS;
        int m = bh;
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) m = imax(m, __shfl_xor_sync(gmask, m, o, G));
        int ec = bh == m ? bcol[h] : 0x7fffffff;
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) ec = imin(ec, __shfl_xor_sync(gmask, ec, o, G));
        const unsigned ball = (__ballot_sync(gmask, bh == m && bcol[h] == ec) >> gshift) & (unsigned)((1ull << G) - 1ull);
        const int winner = __ffs((int)ball) - 1;
        int row = 0;
        if (j == winner && m > 0) {
            for (int i = 0; i < S; ++i)
                if (half_of(snap[h * SP + i], h) - SW_BIAS == m) { row = j * S + i - off[h]; break; }
        }
        row = __shfl_sync(gmask, row, winner, G);
        int end_read = 0;
        if (m == 0) ec = 0;                               // nothing scored: ssw.c keeps end_ref = 0, Hmax = 0
        else end_read = imin(rows[h] - 1, row);
        if (rows[h] > 0 && end_read > rows[h] - 1) end_read = rows[h] - 1;

        const bool wr = live && t < d.n_tasks;            // absent tasks run along (full-mask shuffles below)
        int32_t *f = d.fwd + (wr ? t : 0) * 8;
        if (!REV) {
            // second-best score outside the mask around ref_end1 (ssw.c:537-550)
            int s2 = 0, e2 = 0x7fffffff;
            const int mask_len = d.prm.mask_len >= 0 ? d.prm.mask_len : rows[h] / 2;
            if (mask_len >= 15) {
                const int lo = imax(ec - mask_len, 0), hi = imin(ec + mask_len, cols[h]);
                for (int c = j; c < cols[h]; c += G) {
                    if (c >= lo && c < hi) continue;
                    const int v = half_of(mcol[c], h) & 0xffff;
                    if (v > s2) { s2 = v; e2 = c; }
                }
#pragma unroll
                for (int o = G / 2; o > 0; o >>= 1) {
                    const int os = __shfl_xor_sync(gmask, s2, o, G), oe = __shfl_xor_sync(gmask, e2, o, G);
                    if (os > s2 || (os == s2 && oe < e2)) { s2 = os; e2 = oe; }
                }
                if (s2 == 0) e2 = 0;
            } else { s2 = 0; e2 = -1; }
            if (wr && j == 0 && (f[F_FLAGS] & FL_VALID)) {
                f[F_SCORE1] = m; f[F_REF_END1] = ec; f[F_READ_END1] = end_read;
                f[F_SCORE2] = s2; f[F_REF_END2] = e2;
                const int flag = d.prm.flag;
                const bool do_rev = !(flag == 0 || (flag == 2 && m < d.prm.filters));      // ssw.c:824
                f[F_FLAGS] = FL_VALID | (do_rev ? FL_DO_REV : 0);
            }
        } else if (wr && j == 0 && (f[F_FLAGS] & FL_DO_REV)) {
            const int rb = aux_ref_end[h] - ec, qb = aux_read_end[h] - end_read;       // ssw.c:836-837
            f[F_REF_BEGIN1] = rb; f[F_READ_BEGIN1] = qb;
            const int flag = d.prm.flag, score1 = f[F_SCORE1];
            const bool skip = (7 & flag) == 0 || ((2 & flag) != 0 && score1 < d.prm.filters) ||
                              ((4 & flag) != 0 && (aux_ref_end[h] - rb > d.prm.filterd || aux_read_end[h] - qb > d.prm.filterd));
            if (!skip) f[F_FLAGS] |= FL_DO_CIGAR;
        }
    }
}

// --------------------------------------------------------------------------------------
// banded: one thread per task, literal restatement of banded_sw's arithmetic with the three
// direction codes of a cell packed into one byte.  A task whose band outgrows BWMAX is
// pushed to the overflow list (main pass) or flagged (overflow pass).
// --------------------------------------------------------------------------------------
struct BandDev {
    DevCtx c;
    const salt_win_t *wins;
    size_t n_tasks;
    const int32_t *fwd;
    SswParams prm;
    uint8_t *dirs; size_t slot;          // per-thread direction scratch
    uint32_t *ovf_list; uint32_t *ovf_count; uint32_t ovf_cap;
    const uint32_t *in_list; const uint32_t *in_count;   // list-driven passes: their input
    uint32_t *dp_list; uint32_t *dp_count;               // diagonal pass: tasks that need banded_sw's dynamic program
    uint32_t *wide_list; uint32_t *wide_count;           // narrow pass: tasks it hands to the warp-per-task kernel
    uint32_t *wide_aux;                                  // per entry of wide_list: band to try next | maximum so far << 16
    uint32_t *coop_list; uint32_t *coop_aux; uint32_t *coop_count;   // medium pass: what it hands to the warp-per-task kernel
    uint32_t *next_item;                                 // warp-per-task kernel: next unclaimed entry of its list
    salt_ssw_out_t *out; uint32_t *cigars; int cigar_stride;
};

__device__ __forceinline__ int band_u(int w, int i, int j) { int x = i - w; if (x < 0) x = 0; return j - x + 1; }
__device__ __forceinline__ int band_d(int w, int i, int j) { int x = i - w; if (x < 0) x = 0; return j - x; }

// Traceback from the bottom-right corner (ssw.c:634-716); ops are produced last-first and reversed in place.
// code_at(i, x, state) = direction code 1..5 of cell (row i, band slot x) for the matrix the walk is in
// (state 0 = E, 1 = F, 2 = H).  Returns the cigar's true length, or -3 when the walk leaves the band.
template <class CodeAt>
__device__ __forceinline__ int band_traceback(const CodeAt &code_at, int band, int wd, int readLen, int refLen, uint32_t *cg, int cigar_stride)
{
    int i = readLen - 1, j = refLen - 1, e = 0, l = 0, fop = 0, prev = 0, state = 2;
    bool bad = false;
    auto emit = [&](uint32_t v) { if (l < cigar_stride) cg[l] = v; ++l; };
    while (i > 0) {
        const int x = band_d(band, i, j);
        if (x < 0 || x >= wd || j < 0) { bad = true; break; }
        const int code = code_at(i, x, state);
        switch (code) {
        case 1: --i; --j; state = 2; fop = 0; break;
        case 2: --i; state = 0; fop = 1; break;
        case 3: --i; state = 2; fop = 1; break;
        case 4: --j; state = 1; fop = 2; break;
        case 5: --j; state = 2; fop = 2; break;
        default: bad = true; break;
        }
        if (bad) break;
        if (fop == prev) ++e;
        else { emit((uint32_t)e << 4 | (uint32_t)prev); prev = fop; e = 1; }
    }
    if (bad) return -3;
    if (fop == 0) emit((uint32_t)(e + 1) << 4);
    else { emit((uint32_t)e << 4 | (uint32_t)fop); emit(16u); }
    const int stored = l < cigar_stride ? l : cigar_stride;
    for (int a = 0, b = stored - 1; a < b; ++a, --b) { const uint32_t tmp = cg[a]; cg[a] = cg[b]; cg[b] = tmp; }
    // l > cigar_stride: the row holds the LAST cigar_stride ops of the true cigar (the first ones the traceback
    // produced), reversed into order; the true length tells the caller the row is partial
    return l;
}

// ---- narrow bands: banded_sw with band B <= 3 entirely in registers ---------------------------------------------
// Nearly every rescue window aligns without a gap or with a short one, so banded_sw's first band
// (|refLen - readLen| + 1, ssw.c:845) is 1, 2 or 3 and succeeds.  For those the three band rows hb / eb / hc of the
// reference (2B+3 entries each) are registers with compile-time indices, a row's reference symbols are one funnel
// shift, and a row's direction codes are one 32-bit word (four bits per cell).  Indices follow the reference:
// cell (i, j) has slot u = j - max(i-B, 0) + 1, its upper neighbour slot u + D with D = [i > B], its left
// neighbour u - 1.  Needs refLen >= 2B+2 so that no early row is cut short by the window's end (the reference's
// `edge` is then i+B+1 for rows 0..B and 2B+2 afterwards); everything else goes to the serial kernel.
template <int B>
struct NarrowBand {
    static constexpr int W = 2 * B + 3;
    int hb[W], eb[W], hc[W];
    int max;
};

// One row.  D = [i > B] (slot shift against the previous row); FULL: the row holds all 2B+1 cells (no window end in
// reach), so nothing in it is predicated.
// DW: the word holding a row's 2B+1 symbols / direction codes (four bits each): uint32_t up to B = 3, uint64_t up to B = 7.
template <int B, int D, bool FULL, class DW>
__device__ __forceinline__ void narrow_row(NarrowBand<B> &s, const int8_t *s_tab, int i, int refLen,
                                           DW symw, int rc, int gapO, int gapE, DW *rowdirs)
{
    constexpr int W = NarrowBand<B>::W;
    const int beg = D ? i - B : 0;
    int n = 2 * B + 1;
    if (!FULL) {
        int end = i + B;
        if (end > refLen - 1) end = refLen - 1;
        n = end - beg + 1;
        if (!D) {
            const int edge = end + 1 < W - 1 ? end + 1 : W - 1;
#pragma unroll
            for (int k = 1; k < W; ++k) if (k == edge) { s.hb[k] = 0; s.eb[k] = 0; }
        }
    }
    s.hb[0] = s.eb[0] = s.hc[0] = 0;
    if (D) { s.hb[W - 1] = 0; s.eb[W - 1] = 0; }
    const int8_t *__restrict__ srow = s_tab + rc;
    int fcur = 0;
    DW dirw = 0;
#pragma unroll
    for (int u = 1; u <= 2 * B + 1; ++u) {
        if (FULL || u <= n) {
            const int ue = u + D, ud = ue - 1, ub = u - 1;
            int t1 = (!D && i == 0) ? -gapO : s.hb[ue] - gapO;
            int t2 = (!D && i == 0) ? -gapE : s.eb[ue] - gapE;
            const uint32_t e_from_h = t1 > t2 ? 1u : 0u;
            const int ev = t1 > t2 ? t1 : t2;
            s.eb[u] = ev;
            t1 = s.hc[ub] - gapO;
            t2 = fcur - gapE;
            const uint32_t f_from_h = t1 > t2 ? 1u : 0u;
            fcur = t1 > t2 ? t1 : t2;
            const int e1 = ev > 0 ? ev : 0;
            const int f1 = fcur > 0 ? fcur : 0;
            t1 = e1 > f1 ? e1 : f1;
            const int sym = (int)((symw >> (4 * (u - 1))) & 15u);
            t2 = s.hb[ud] + srow[sym * 8];
            const int h = t1 > t2 ? t1 : t2;
            s.hc[u] = h;
            s.max = h > s.max ? h : s.max;
            const uint32_t hsel = t1 <= t2 ? 0u : (e1 > f1 ? 1u : 2u);
            dirw |= (DW)(e_from_h | (f_from_h << 1) | (hsel << 2)) << (4 * (u - 1));
        }
    }
#pragma unroll
    for (int k = 1; k <= 2 * B + 1; ++k) if (FULL || k <= n) s.hb[k] = s.hc[k];
    rowdirs[(size_t)i * 32] = dirw;
}

// one banded_sw attempt at band B; max carries over between attempts as in the reference's do-while (ssw.c:575-632)
template <int B, bool PAC, class DW>
__device__ __forceinline__ bool narrow_fill(const DevCtx &c, const int8_t *s_tab, uint32_t rs, int read0, uint32_t ref0,
                                            int readLen, int refLen, int score, int gapO, int gapE, DW *rowdirs, int &max)
{
    static_assert(4 * (2 * B + 1) <= 8 * (int)sizeof(DW), "a row's codes must fit the word");
    NarrowBand<B> s;
#pragma unroll
    for (int k = 0; k < NarrowBand<B>::W; ++k) { s.hb[k] = 0; s.eb[k] = 0; s.hc[k] = 0; }
    s.max = max;
    // a row's 2B+1 reference symbols are one nibble word; it and the read's code word are requested a row / a word ahead
    auto row_syms = [&](int i, int D) -> DW {
        const uint32_t p = ref0 + (uint32_t)(D ? i - B : 0);
        if (PAC) {
            uint64_t v = sw_pac_group(c.pac, p);
            if (sizeof(DW) > 4) v |= (uint64_t)sw_pac_group(c.pac, p + 8u) << 32;
            return (DW)v;
        }
        const uint32_t *__restrict__ mw = c.mixref + (p >> 3);
        const int sh = 4 * (int)(p & 7u);
        const uint32_t w1 = mw[1];
        uint64_t v = __funnelshift_r(mw[0], w1, sh);
        if (sizeof(DW) > 4) v |= (uint64_t)__funnelshift_r(w1, mw[2], sh) << 32;
        return (DW)v;
    };
    const uint64_t *__restrict__ rd = c.rd4 + (size_t)rs * c.W64;
    const int W64 = (int)c.W64;
    uint64_t rw = rd[read0 >> 4], rwn = (read0 >> 4) + 1 < W64 ? rd[(read0 >> 4) + 1] : 0ull;
    bool first_code = true;
    auto code_of = [&](int idx) {
        if ((idx & 15) == 0 && !first_code) { rw = rwn; rwn = (idx >> 4) + 1 < W64 ? rd[(idx >> 4) + 1] : 0ull; }
        first_code = false;
        const unsigned nib = (unsigned)(rw >> (4 * (idx & 15))) & 15u;
        return nib == 15u ? SW_CODE_N : (31 - __clz(nib));
    };
#pragma unroll
    for (int i = 0; i <= B; ++i)
        if (i < readLen) narrow_row<B, 0, false, DW>(s, s_tab, i, refLen, row_syms(i, 0), code_of(read0 + i), gapO, gapE, rowdirs);
    int full_end = refLen - B;                                  // rows below it hold all 2B+1 cells: i + B <= refLen - 1
    if (full_end > readLen) full_end = readLen;
    int i = B + 1;
    DW nxt = row_syms(i, 1);
    for (; i < full_end; ++i) {
        const DW cur = nxt;
        nxt = row_syms(i + 1, 1);
        narrow_row<B, 1, true, DW>(s, s_tab, i, refLen, cur, code_of(read0 + i), gapO, gapE, rowdirs);
    }
    for (; i < readLen; ++i) {
        const DW cur = nxt;
        nxt = row_syms(i + 1, 1);
        narrow_row<B, 1, false, DW>(s, s_tab, i, refLen, cur, code_of(read0 + i), gapO, gapE, rowdirs);
    }
    max = s.max;
    return s.max >= score;
}

// direction codes of the narrow pass for the traceback; the walk only moves up one row at a time, so the word of the row
// above is already on its way when it is needed
template <class DW>
struct NarrowCodeAt {
    const DW *rowdirs;
    mutable int ci; mutable DW wc, wp;
    __device__ __forceinline__ int operator()(int i, int x, int state) const
    {
        if (i != ci) {
            wc = i == ci - 1 ? wp : rowdirs[(size_t)i * 32];
            ci = i;
            wp = i > 0 ? rowdirs[(size_t)(i - 1) * 32] : (DW)0;
        }
        const uint32_t nib = (uint32_t)(wc >> (4 * x)) & 15u;
        const int ce = (nib & 1u) ? 3 : 2, cf = (nib & 2u) ? 5 : 4;
        if (state == 0) return ce;
        if (state == 1) return cf;
        const uint32_t hs = nib >> 2;
        return hs == 0u ? 1 : (hs == 1u ? ce : cf);
    }
};

// ---- the gapless majority: no dynamic program at all ----------------------------------------------------------------------
// banded_sw runs on the rectangle [ref_begin1, ref_end1] x [read_begin1, read_end1] the two score passes found.  When the
// rectangle is square and the plain sum P of the substitution scores along its diagonal equals score1, the result is known:
// the first band is 1; H(i,i) >= P_i along the diagonal, so the band's maximum reaches score1 and the band is not doubled;
// and at every diagonal cell the traceback takes the diagonal -- were max(E, F, 0) > H(i-1,i-1) + s_i >= P_i at some
// cell, following the diagonal from there would end above P = score1, the maximum over the whole window (ties go to the
// diagonal, ssw.c:609).  The cigar is "<readLen>M".  Everything else is listed for the dynamic program.
template <bool PAC>
__global__ void __launch_bounds__(128)
sw_diag_kernel(BandDev d)
{
    __shared__ int8_t s_tab[17 * 8];
    for (int i = threadIdx.x; i < 17 * 8; i += blockDim.x) s_tab[i] = d.prm.table[i];
    __syncthreads();
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= d.n_tasks) return;
    const int32_t *f = d.fwd + t * 8;
    const int fl = f[F_FLAGS];
    salt_ssw_out_t o;
    o.score1 = (uint16_t)f[F_SCORE1]; o.score2 = (uint16_t)f[F_SCORE2];
    o.ref_begin1 = f[F_REF_BEGIN1]; o.ref_end1 = f[F_REF_END1];
    o.read_begin1 = f[F_READ_BEGIN1]; o.read_end1 = f[F_READ_END1];
    o.ref_end2 = f[F_REF_END2]; o.cigarLen = 0;
    if (!(fl & FL_VALID)) { o.ref_end2 = -1; o.cigarLen = -1; }
    if (!(fl & FL_DO_CIGAR)) { d.out[t] = o; return; }
    const salt_win_t w = d.wins[t];
    const int refLen = o.ref_end1 - o.ref_begin1 + 1, readLen = o.read_end1 - o.read_begin1 + 1;
    if (refLen == readLen && readLen >= 1) {
        const uint32_t ref0 = w.start + (uint32_t)o.ref_begin1;
        const int read0 = o.read_begin1;
        const uint64_t *__restrict__ rd = d.c.rd4 + (size_t)w.rs * d.c.W64;
        int P = 0;
        for (int i0 = 0; i0 < readLen; i0 += 8) {
            const uint32_t p = ref0 + (uint32_t)i0;
            uint32_t symw;
            if (PAC) symw = sw_pac_group(d.c.pac, p);
            else { const uint32_t *__restrict__ mw = d.c.mixref + (p >> 3); symw = __funnelshift_r(mw[0], mw[1], 4 * (int)(p & 7u)); }
            const int idx = read0 + i0, wi = idx >> 4, sh = 4 * (idx & 15);
            uint64_t rw = rd[wi] >> sh;
            if (sh > 32 && wi + 1 < (int)d.c.W64) rw |= rd[wi + 1] << (64 - sh);
            const uint32_t codes = (uint32_t)rw;
            const int n8 = readLen - i0 < 8 ? readLen - i0 : 8;
#pragma unroll
            for (int m = 0; m < 8; ++m) {
                if (m < n8) {
                    const unsigned nib = (codes >> (4 * m)) & 15u;
                    const int rc = nib == 15u ? SW_CODE_N : (31 - __clz(nib));
                    P += s_tab[((symw >> (4 * m)) & 15u) * 8 + rc];
                }
            }
        }
        if (P == (int)o.score1) {
            if (d.cigar_stride > 0) d.cigars[t * (size_t)d.cigar_stride] = (uint32_t)readLen << 4;
            o.cigarLen = 1;
            d.out[t] = o;
            return;
        }
    }
    d.dp_list[atomicAdd(d.dp_count, 1u)] = (uint32_t)t;
}

template <bool PAC>
__global__ void __launch_bounds__(128)
sw_banded_narrow_kernel(BandDev d)
{
    __shared__ int8_t s_tab[17 * 8];
    for (int i = threadIdx.x; i < 17 * 8; i += blockDim.x) s_tab[i] = d.prm.table[i];
    __syncthreads();
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t n_items = (size_t)*d.in_count;
    const size_t item_step = (size_t)gridDim.x * blockDim.x;
    for (size_t item = tid; item < n_items; item += item_step) {
    const size_t t = (size_t)d.in_list[item];
    const int32_t *f = d.fwd + t * 8;
    salt_ssw_out_t o;
    o.score1 = (uint16_t)f[F_SCORE1]; o.score2 = (uint16_t)f[F_SCORE2];
    o.ref_begin1 = f[F_REF_BEGIN1]; o.ref_end1 = f[F_REF_END1];
    o.read_begin1 = f[F_READ_BEGIN1]; o.read_end1 = f[F_READ_END1];
    o.ref_end2 = f[F_REF_END2]; o.cigarLen = 0;                  // listed tasks are valid and want a cigar

    const salt_win_t w = d.wins[t];
    const int refLen = o.ref_end1 - o.ref_begin1 + 1, readLen = o.read_end1 - o.read_begin1 + 1;
    const uint32_t ref0 = w.start + (uint32_t)o.ref_begin1;
    const int read0 = o.read_begin1;
    const int score = o.score1, gapO = d.prm.gapO, gapE = d.prm.gapE;
    int band = abs(refLen - readLen) + 1;                                    // ssw.c:845
    // a row's direction word of the 32 tasks of a warp side by side: [row][lane]
    uint32_t *rowdirs = reinterpret_cast<uint32_t *>(d.dirs + (tid >> 5) * (d.slot * 32)) + (tid & 31);
    bool ok = false;
    int max = 0;
    if (readLen >= 1 && (size_t)readLen * 4 <= d.slot) {
        if (band == 1 && refLen >= 4) {
            ok = narrow_fill<1, PAC, uint32_t>(d.c, s_tab, w.rs, read0, ref0, readLen, refLen, score, gapO, gapE, rowdirs, max);
            if (!ok) band = 2;
        }
        if (!ok && band == 2 && refLen >= 6) {
            ok = narrow_fill<2, PAC, uint32_t>(d.c, s_tab, w.rs, read0, ref0, readLen, refLen, score, gapO, gapE, rowdirs, max);
            if (!ok) band = 4;
        } else if (!ok && band == 3 && refLen >= 8) {
            ok = narrow_fill<3, PAC, uint32_t>(d.c, s_tab, w.rs, read0, ref0, readLen, refLen, score, gapO, gapE, rowdirs, max);
            if (!ok) band = 6;
        }
    }
    if (!ok) {                               // the warp-per-task kernel goes on where this one stopped: next band, maximum so far
        const uint32_t k = atomicAdd(d.wide_count, 1u);
        d.wide_list[k] = (uint32_t)t;
        d.wide_aux[k] = (uint32_t)(band < 0xffff ? band : 0xffff) | ((uint32_t)max << 16);
        continue;
    }
    NarrowCodeAt<uint32_t> at{rowdirs, -1000, 0u, 0u};
    o.cigarLen = band_traceback(at, band, 2 * band + 1, readLen, refLen, d.cigars + t * (size_t)d.cigar_stride, d.cigar_stride);
    d.out[t] = o;
    }
}

// ---- bands 4..7: still one thread per task and registers only (a row's fifteen codes are one 64-bit word) ---------------
// Indel-rich reads (several gaps per read, BASELINE configs[4]) put most rescue windows here; a warp per task would spend
// a hundred instructions per step on nine to fifteen lanes.  Walks the list the narrow pass left (task, band to try,
// maximum so far), makes ONE attempt at that band and either finishes the task or passes it on with the doubled band.
template <bool PAC>
__global__ void __launch_bounds__(128)
sw_banded_medium_kernel(BandDev d)
{
    __shared__ int8_t s_tab[17 * 8];
    for (int i = threadIdx.x; i < 17 * 8; i += blockDim.x) s_tab[i] = d.prm.table[i];
    __syncthreads();
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t n_items = (size_t)*d.in_count;
    const size_t item_step = (size_t)gridDim.x * blockDim.x;
    for (size_t item = tid; item < n_items; item += item_step) {
        const size_t t = (size_t)d.in_list[item];
        const uint32_t aux = d.wide_aux[item];
        int band = (int)(aux & 0xffffu), max = (int)(aux >> 16);
        const int32_t *f = d.fwd + t * 8;
        salt_ssw_out_t o;
        o.score1 = (uint16_t)f[F_SCORE1]; o.score2 = (uint16_t)f[F_SCORE2];
        o.ref_begin1 = f[F_REF_BEGIN1]; o.ref_end1 = f[F_REF_END1];
        o.read_begin1 = f[F_READ_BEGIN1]; o.read_end1 = f[F_READ_END1];
        o.ref_end2 = f[F_REF_END2]; o.cigarLen = 0;
        const salt_win_t w = d.wins[t];
        const int refLen = o.ref_end1 - o.ref_begin1 + 1, readLen = o.read_end1 - o.read_begin1 + 1;
        const uint32_t ref0 = w.start + (uint32_t)o.ref_begin1;
        const int read0 = o.read_begin1;
        const int score = o.score1, gapO = d.prm.gapO, gapE = d.prm.gapE;
        uint64_t *rowdirs = reinterpret_cast<uint64_t *>(d.dirs + (tid >> 5) * (d.slot * 32)) + (tid & 31);
        bool ok = false;
        if (band >= 4 && band <= 7 && readLen >= 1 && refLen >= 2 * band + 2 && (size_t)readLen * 8 <= d.slot) {
            const int b0 = band;
            switch (b0) {
            case 4: ok = narrow_fill<4, PAC, uint64_t>(d.c, s_tab, w.rs, read0, ref0, readLen, refLen, score, gapO, gapE, rowdirs, max); break;
            case 5: ok = narrow_fill<5, PAC, uint64_t>(d.c, s_tab, w.rs, read0, ref0, readLen, refLen, score, gapO, gapE, rowdirs, max); break;
            case 6: ok = narrow_fill<6, PAC, uint64_t>(d.c, s_tab, w.rs, read0, ref0, readLen, refLen, score, gapO, gapE, rowdirs, max); break;
            default: ok = narrow_fill<7, PAC, uint64_t>(d.c, s_tab, w.rs, read0, ref0, readLen, refLen, score, gapO, gapE, rowdirs, max); break;
            }
            if (!ok) band = 2 * b0;
        }
        if (!ok) {
            const uint32_t k = atomicAdd(d.coop_count, 1u);
            d.coop_list[k] = (uint32_t)t;
            d.coop_aux[k] = (uint32_t)(band < 0xffff ? band : 0xffff) | ((uint32_t)max << 16);
            continue;
        }
        NarrowCodeAt<uint64_t> at{rowdirs, -1000, 0ull, 0ull};
        o.cigarLen = band_traceback(at, band, 2 * band + 1, readLen, refLen, d.cigars + t * (size_t)d.cigar_stride, d.cigar_stride);
        d.out[t] = o;
    }
}

// ---- wide bands: one warp per task, lanes over the band's diagonals ----------------------------------------------------
// A task the narrow pass hands over has a first band of 4+ (a long gap) or needed its band doubled: a few per thousand,
// but one thread walking (2B+1) x readLen cells after the other is what the whole stage then waits for.  Here lane k owns
// diagonal j - i + B = k of the band.  Cell (i, k) needs (i, k-1) [left: H, F], (i-1, k+1) [up: H, E] and (i-1, k) [own
// previous cell: H], so with lane k working on row i at step s = 2i + k both neighbours finished exactly one step earlier
// and their values arrive by shuffle; a lane whose cell lies outside the window or the read publishes zeros, which is what
// banded_sw's zeroed edge slots hold (refLen >= 2B+2, as for the narrow pass; anything else goes to the serial kernel).
constexpr int COOP_MAXB = 15;                    // 2B+1 <= 31 lanes

struct PlainCodeAt {
    const uint8_t *dirs; int wd;
    __device__ __forceinline__ int operator()(int i, int x, int state) const { return sw_dir_get(dirs[(size_t)wd * i + x], state); }
};

__global__ void __launch_bounds__(128)
sw_banded_coop_kernel(BandDev d, int rows8, int smem_per_warp)
{
    constexpr unsigned FULLM = 0xffffffffu;
    __shared__ int8_t s_tab[17 * 8];
    SALT_DYN_SMEM(uint8_t, s_dirs);                                      // [warp][smem_per_warp]
    for (int i = threadIdx.x; i < 17 * 8; i += blockDim.x) s_tab[i] = d.prm.table[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const size_t n_items = (size_t)*d.in_count;
    uint8_t *gdirs = d.dirs + warp * ((size_t)(2 * COOP_MAXB + 1) * (size_t)rows8);
    uint8_t *sdirs = s_dirs + (size_t)(threadIdx.x >> 5) * (size_t)smem_per_warp;
    for (;;) {
        // tasks differ by an order of magnitude (one attempt at band 4 ... four attempts up to band 8): warps take the next
        // listed task when they are free instead of a fixed share
        uint32_t item = 0;
        if (lane == 0) item = atomicAdd(d.next_item, 1u);
        item = __shfl_sync(FULLM, item, 0);
        if ((size_t)item >= n_items) break;
        const size_t t = (size_t)d.in_list[item];
        const int32_t *f = d.fwd + t * 8;
        salt_ssw_out_t o;
        o.score1 = (uint16_t)f[F_SCORE1]; o.score2 = (uint16_t)f[F_SCORE2];
        o.ref_begin1 = f[F_REF_BEGIN1]; o.ref_end1 = f[F_REF_END1];
        o.read_begin1 = f[F_READ_BEGIN1]; o.read_end1 = f[F_READ_END1];
        o.ref_end2 = f[F_REF_END2]; o.cigarLen = 0;
        const salt_win_t w = d.wins[t];
        const int refLen = o.ref_end1 - o.ref_begin1 + 1, readLen = o.read_end1 - o.read_begin1 + 1;
        const uint32_t ref0 = w.start + (uint32_t)o.ref_begin1;
        const int read0 = o.read_begin1;
        const int score = o.score1, gapO = d.prm.gapO, gapE = d.prm.gapE;
        const uint32_t aux = d.coop_aux[item];                               // where the passes before stopped (ssw.c:845 and the doublings they tried)
        int band = (int)(aux & 0xffffu);
        int max = (int)(aux >> 16), wd = 0;
        bool served = true;
        uint8_t *dirs = gdirs;
        do {
            wd = 2 * band + 1;
            if (band > COOP_MAXB || refLen < 2 * band + 2 || readLen < 1) { served = false; break; }
            dirs = (size_t)wd * (size_t)readLen <= (size_t)smem_per_warp ? sdirs : gdirs;
            const int k = lane;
            const bool lane_on = k < wd;
            int Hlast = 0, Elast = 0, Flast = 0, Hown = 0, lmax = 0;
            // The lane walks its diagonal: row and column advance together.  Eight reference symbols and sixteen read codes
            // at a time sit in registers, the words behind them are fetched a group ahead, so no step waits on memory.
            const int i0 = band - k > 0 ? band - k : 0;                    // first row whose column j = i + k - band is >= 0
            const uint32_t p0 = ref0 + (uint32_t)(i0 + k - band);
            uint32_t symw = 0, w1 = 0, w2 = 0, wnext = 0;
            int rsh = 0;
            if (!d.prm.use_pac) {
                const uint32_t *__restrict__ mw = d.c.mixref + (p0 >> 3);
                rsh = 4 * (int)(p0 & 7u);
                w1 = mw[1]; w2 = mw[2]; wnext = 3;
                symw = __funnelshift_r(mw[0], w1, rsh);
            } else {
                symw = sw_pac_group(d.c.pac, p0); w1 = sw_pac_group(d.c.pac, p0 + 8u);
            }
            const uint64_t *__restrict__ rd = d.c.rd4 + (size_t)w.rs * d.c.W64;
            const int rword0 = (read0 + i0) >> 4;
            uint64_t rw = rword0 < (int)d.c.W64 ? rd[rword0] : 0ull, rwn = rword0 + 1 < (int)d.c.W64 ? rd[rword0 + 1] : 0ull;
            const int n_steps = 2 * (readLen - 1) + wd;
            for (int s = 0; s < n_steps; ++s) {
                const int upH = __shfl_down_sync(FULLM, Hlast, 1), upE = __shfl_down_sync(FULLM, Elast, 1);
                int leftH = __shfl_up_sync(FULLM, Hlast, 1), leftF = __shfl_up_sync(FULLM, Flast, 1);
                if (lane == 0) { leftH = 0; leftF = 0; }
                const int d2 = s - k;
                if (lane_on && d2 >= 0 && !(d2 & 1)) {
                    const int i = d2 >> 1, j = i + k - band;
                    int h = 0, e = 0, fv = 0;
                    if (i >= i0) {
                        const int cell = i - i0;                            // cells of this lane so far
                        const int sym = (int)((symw >> (4 * (cell & 7))) & 15u);
                        const int ridx = read0 + i;
                        const unsigned nib = (unsigned)(rw >> (4 * (ridx & 15))) & 15u;
                        const int rc = nib == 15u ? SW_CODE_N : (31 - __clz(nib));
                        if (i < readLen && j < refLen) {
                            int t1 = upH - gapO, t2 = upE - gapE;
                            const int de = t1 > t2 ? 3 : 2;
                            e = t1 > t2 ? t1 : t2;
                            t1 = leftH - gapO; t2 = leftF - gapE;
                            const int df = t1 > t2 ? 5 : 4;
                            fv = t1 > t2 ? t1 : t2;
                            const int e1 = e > 0 ? e : 0, f1 = fv > 0 ? fv : 0;
                            t1 = e1 > f1 ? e1 : f1;
                            t2 = Hown + s_tab[sym * 8 + rc];
                            h = t1 > t2 ? t1 : t2;
                            lmax = h > lmax ? h : lmax;
                            const int dh = t1 <= t2 ? 1 : (e1 > f1 ? de : df);
                            dirs[(size_t)wd * i + band_d(band, i, j)] = sw_dir_pack(de, df, dh);
                        }
                        if ((cell & 7) == 7) {                              // next eight symbols; the words after them are requested now
                            if (!d.prm.use_pac) {
                                symw = __funnelshift_r(w1, w2, rsh);
                                w1 = w2; w2 = d.c.mixref[(p0 >> 3) + wnext]; ++wnext;
                            } else {
                                symw = w1; w1 = sw_pac_group(d.c.pac, p0 + (uint32_t)(cell + 9));
                            }
                        }
                        if ((ridx & 15) == 15) {
                            rw = rwn;
                            const int nw = (ridx >> 4) + 2;
                            rwn = nw < (int)d.c.W64 ? rd[nw] : 0ull;
                        }
                    }
                    Hown = h; Hlast = h; Elast = e; Flast = fv;
                }
            }
#pragma unroll
            for (int o2 = 16; o2 > 0; o2 >>= 1) { const int v = __shfl_xor_sync(FULLM, lmax, o2); lmax = v > lmax ? v : lmax; }
            max = lmax > max ? lmax : max;
            band *= 2;
        } while (max < score);
        __syncwarp();
        if (lane == 0) {
            if (!served) {
                if (d.ovf_list) { d.ovf_list[atomicAdd(d.ovf_count, 1u)] = (uint32_t)t; o.cigarLen = 0; }
                else o.cigarLen = -2;
            } else {
                band /= 2;
                PlainCodeAt at{dirs, wd};
                o.cigarLen = band_traceback(at, band, wd, readLen, refLen, d.cigars + t * (size_t)d.cigar_stride, d.cigar_stride);
            }
            d.out[t] = o;
        }
        __syncwarp();
    }
}

// The serial kernel: one thread per listed task, a literal restatement of banded_sw's arrays (any window shape, every quirk
// of its edge handling).  It serves what the two passes above hand on (FINAL = true: SSW_OVF_THREADS threads with scratch
// for bands up to BWMAX = 512, a band beyond that is declined per item); every thread strides over the list.
template <int BWMAX, bool FINAL>
__global__ void __launch_bounds__(128)
sw_banded_kernel(BandDev d)
{
    __shared__ int8_t s_tab[17 * 8];
    for (int i = threadIdx.x; i < 17 * 8; i += blockDim.x) s_tab[i] = d.prm.table[i];
    __syncthreads();
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t n_items = (size_t)*d.in_count;
    const size_t item_step = (size_t)gridDim.x * blockDim.x;
    for (size_t item = tid; item < n_items; item += item_step) {
    const size_t t = (size_t)d.in_list[item];
    const int32_t *f = d.fwd + t * 8;
    salt_ssw_out_t o;
    o.score1 = (uint16_t)f[F_SCORE1]; o.score2 = (uint16_t)f[F_SCORE2];
    o.ref_begin1 = f[F_REF_BEGIN1]; o.ref_end1 = f[F_REF_END1];
    o.read_begin1 = f[F_READ_BEGIN1]; o.read_end1 = f[F_READ_END1];
    o.ref_end2 = f[F_REF_END2]; o.cigarLen = 0;                  // listed tasks are valid and want a cigar

    const salt_win_t w = d.wins[t];
    const int refLen = o.ref_end1 - o.ref_begin1 + 1, readLen = o.read_end1 - o.read_begin1 + 1;
    const uint32_t ref0 = w.start + (uint32_t)o.ref_begin1;
    const int read0 = o.read_begin1;
    const int score = o.score1, gapO = d.prm.gapO, gapE = d.prm.gapE;
    int band = abs(refLen - readLen) + 1;                                    // ssw.c:845
    // direction bytes of the 32 tasks of a warp are interleaved ([cell][lane]): the tasks of a warp walk
    // their bands in step, so a warp's byte stores and traceback loads land in one 32-byte sector
    uint8_t *dirs = d.dirs + (tid >> 5) * (d.slot * 32) + (tid & 31);
    int hb[2 * BWMAX + 5], eb[2 * BWMAX + 5], hc[2 * BWMAX + 5];
    int max = 0, wd = 0;
    bool overflow = false;
    do {
        const int width = band * 2 + 3;
        wd = band * 2 + 1;
        if (band > BWMAX || (size_t)wd * (size_t)readLen > d.slot) { overflow = true; break; }
        for (int j = 1; j < width - 1; ++j) hb[j] = 0;
        for (int i = 0; i < readLen; ++i) {
            int beg = 0, end = refLen - 1, u = 0;
            if (i - band > beg) beg = i - band;
            if (i + band < end) end = i + band;
            const int edge = end + 1 < width - 1 ? end + 1 : width - 1;
            int fcur = 0;
            hb[0] = eb[0] = hb[edge] = eb[edge] = hc[0] = 0;
            uint8_t *dl = dirs + (size_t)wd * i * 32;
            const int rc = sw_read_code(d.c, w.rs, read0 + i);
            for (int j = beg; j <= end; ++j) {
                u = band_u(band, i, j);
                const int ue = band_u(band, i - 1, j), ub = band_u(band, i, j - 1), ud = band_u(band, i - 1, j - 1);
                int t1 = i == 0 ? -gapO : hb[ue] - gapO;
                int t2 = i == 0 ? -gapE : eb[ue] - gapE;
                eb[u] = t1 > t2 ? t1 : t2;
                const int de = t1 > t2 ? 3 : 2;
                t1 = hc[ub] - gapO;
                t2 = fcur - gapE;
                fcur = t1 > t2 ? t1 : t2;
                const int df = t1 > t2 ? 5 : 4;
                const int e1 = eb[u] > 0 ? eb[u] : 0;
                const int f1 = fcur > 0 ? fcur : 0;
                t1 = e1 > f1 ? e1 : f1;
                const int sym = sw_ref_symbol(d.c, d.prm.use_pac, ref0 + (uint32_t)j);
                t2 = hb[ud] + s_tab[sym * 8 + rc];
                hc[u] = t1 > t2 ? t1 : t2;
                if (hc[u] > max) max = hc[u];
                const int dh = t1 <= t2 ? 1 : (e1 > f1 ? de : df);
                dl[band_d(band, i, j) * 32] = sw_dir_pack(de, df, dh);
            }
            for (int j = 1; j <= u; ++j) hb[j] = hc[j];
        }
        band *= 2;
    } while (max < score);

    if (overflow) {
        if (!FINAL && d.ovf_list) {
            const uint32_t k = atomicAdd(d.ovf_count, 1u);     // the list has room for every task
            d.ovf_list[k] = (uint32_t)t; o.cigarLen = 0;
        } else o.cigarLen = -2;                                // band beyond the engine's widest (BWMAX of the overflow pass)
        d.out[t] = o;
        continue;
    }
    band /= 2;

    struct ByteCodeAt {
        const uint8_t *dirs; int wd;
        __device__ __forceinline__ int operator()(int i, int x, int state) const { return sw_dir_get(dirs[((size_t)wd * i + x) * 32], state); }
    };
    ByteCodeAt at{dirs, wd};
    o.cigarLen = band_traceback(at, band, wd, readLen, refLen, d.cigars + t * (size_t)d.cigar_stride, d.cigar_stride);
    d.out[t] = o;
    }
}

// --------------------------------------------------------------------------------------
// host side of this file
// --------------------------------------------------------------------------------------
struct SwShape { int G, S; };

// Strip shapes.  Fewer, taller strips amortise the per-column work (score lookup, three shuffles,
// best tracking) over more rows and shorten the systolic fill; 4 threads x 26 rows serves the
// 100 bp case (104 padded rows), 8 threads per pair the longer reads.
static SwShape pick_shape(int l_max)
{
    const int seg = (l_max + 7) / 8;                 // rows needed = 8*seg
    static const int s4[] = {16, 26};
    if (!getenv("SALT_SW_G8"))
        for (int s : s4) if (2 * seg <= s) return {4, s};
    static const int s8[] = {8, 13, 16, 19, 24, 32};
    for (int s : s8) if (seg <= s) return {8, s};
    if (8 * seg <= 16 * 32) return {16, 32};
    return {32, 32};
}

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// layout: [0]=win2 [1]=unused [2]=maxcol2 [3]=fwd [4]=dirs [5]=ovf_list [6]=counters [7]=total [8]=wide_list [9]=dp_list [10]=wide_aux [11]=coop_list [12]=coop_aux
size_t ssw_scratch_bytes(size_t n_tasks, int max_cols, int max_rows, size_t *layout)
{
    const size_t pairs = (n_tasks + 1) / 2;
    const int CW = (max_cols + 7) / 8 + 1;
    size_t off = 0;
    layout[0] = off; off = align_up(off + pairs * CW * sizeof(uint2), 256);
    layout[1] = off;                                  // (selectors: made inside the DP kernel since round 2)
    layout[2] = off; off = align_up(off + pairs * (size_t)max_cols * sizeof(uint32_t), 256);
    layout[3] = off; off = align_up(off + pairs * 2 * 8 * sizeof(int32_t), 256);
    const size_t slot = (size_t)(2 * 16 + 1) * (size_t)(8 * ((max_rows + 7) / 8));
    layout[4] = off; off = align_up(off + align_up(n_tasks, 128) * slot, 256);
    layout[5] = off; off = align_up(off + n_tasks * sizeof(uint32_t), 256);
    layout[6] = off; off = align_up(off + 256, 256);
    layout[8] = off; off = align_up(off + n_tasks * sizeof(uint32_t), 256);
    layout[9] = off; off = align_up(off + n_tasks * sizeof(uint32_t), 256);
    layout[10] = off; off = align_up(off + n_tasks * sizeof(uint32_t), 256);
    layout[11] = off; off = align_up(off + n_tasks * sizeof(uint32_t), 256);
    layout[12] = off; off = align_up(off + n_tasks * sizeof(uint32_t), 256);
    layout[7] = off;
    return off;
}

template <int G, int S>
static cudaError_t run_dp(const SwDev &d, bool rev, cudaStream_t st)
{
    constexpr int GPC = 128 / G;
    const size_t blocks = (d.n_pairs + GPC - 1) / GPC;
    const size_t smem = 128 * 2 * ((S + 3) / 4 * 4) * sizeof(uint32_t);
    if (rev) {
        auto kern = sw_dp_kernel<G, S, true>;
        if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        SALT_LAUNCH(kern, (unsigned)blocks, 128, smem, st, d);
    } else {
        auto kern = sw_dp_kernel<G, S, false>;
        if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        SALT_LAUNCH(kern, (unsigned)blocks, 128, smem, st, d);
    }
    return cudaGetLastError();
}

static cudaError_t dispatch_dp(const SwDev &d, bool rev, cudaStream_t st)
{
    switch (d.G * 100 + d.S) {
    case 416: return run_dp<4, 16>(d, rev, st);
    case 426: return run_dp<4, 26>(d, rev, st);
    case 808: return run_dp<8, 8>(d, rev, st);
    case 813: return run_dp<8, 13>(d, rev, st);
    case 816: return run_dp<8, 16>(d, rev, st);
    case 819: return run_dp<8, 19>(d, rev, st);
    case 824: return run_dp<8, 24>(d, rev, st);
    case 832: return run_dp<8, 32>(d, rev, st);
    case 1632: return run_dp<16, 32>(d, rev, st);
    case 3232: return run_dp<32, 32>(d, rev, st);
    }
    return cudaErrorInvalidValue;
}

// Overflow pass for bands wider than 16: SSW_OVF_THREADS threads, each with direction scratch for band <= 512
// over the chunk's longest read, walk the overflow list however long it is.  The scratch belongs to the handle.
size_t ssw_overflow_bytes(int max_rows)
{
    return (size_t)SSW_OVF_THREADS * (size_t)(2 * 512 + 1) * (size_t)(8 * ((max_rows + 7) / 8));
}

cudaError_t launch_ssw(const DevCtx &c, const salt_win_t *wins, size_t n, const SswParams &prm,
                       void *scratch, size_t scratch_bytes, int max_cols,
                       salt_ssw_out_t *out, uint32_t *cigars, int cigar_stride, int sm_count, cudaStream_t st,
                       uint64_t *launches, cudaEvent_t *ev, uint8_t *ovf_dirs)
{
#define SALT_EV(i) do { if (ev) cudaEventRecord(ev[i], st); } while (0)
    if (!n) return cudaSuccess;
    size_t lay[13];
    const size_t need = ssw_scratch_bytes(n, max_cols, (int)c.l_max, lay);
    if (need > scratch_bytes) return cudaErrorMemoryAllocation;
    const SwShape sh = pick_shape((int)c.l_max);
    uint8_t *base = static_cast<uint8_t *>(scratch);
    SwDev d;
    d.c = c; d.wins = wins; d.n_tasks = n; d.n_pairs = (n + 1) / 2;
    d.G = sh.G; d.S = sh.S; d.CW = (max_cols + 7) / 8 + 1; d.MC = max_cols;
    d.win2 = reinterpret_cast<uint2 *>(base + lay[0]);
    d.maxcol2 = reinterpret_cast<uint32_t *>(base + lay[2]);
    d.fwd = reinterpret_cast<int32_t *>(base + lay[3]);
    d.prm = prm;
    d.item_shift = 0;
    while ((1 << d.item_shift) < d.CW) ++d.item_shift;
    const size_t prep_items = (d.n_pairs * 2) << d.item_shift;
    const unsigned prep_blocks = (unsigned)((prep_items + 255) / 256);
    cudaError_t e;
    uint32_t *ovf_count = reinterpret_cast<uint32_t *>(base + lay[6]);
    if ((e = cudaMemsetAsync(ovf_count, 0, 256, st)) != cudaSuccess) return e;

    SALT_EV(0);
    { auto kern = sw_prep_kernel<false>; SALT_LAUNCH(kern, prep_blocks, 256, 0, st, d); }
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    SALT_EV(1);
    if ((e = dispatch_dp(d, false, st)) != cudaSuccess) return e;
    SALT_EV(2);
    { auto kern = sw_prep_kernel<true>; SALT_LAUNCH(kern, prep_blocks, 256, 0, st, d); }
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    SALT_EV(3);
    if ((e = dispatch_dp(d, true, st)) != cudaSuccess) return e;
    SALT_EV(4);

    BandDev b;
    b.c = c; b.wins = wins; b.n_tasks = n; b.fwd = d.fwd; b.prm = prm;
    b.dirs = base + lay[4];
    b.slot = (size_t)(2 * 16 + 1) * (size_t)(8 * (((int)c.l_max + 7) / 8));
    b.ovf_list = reinterpret_cast<uint32_t *>(base + lay[5]);
    b.ovf_count = ovf_count; b.ovf_cap = SSW_OVF_THREADS;
    if (!ovf_dirs) b.ovf_list = nullptr;                  // no overflow scratch: wide bands come back with cigarLen = -2
    b.wide_list = reinterpret_cast<uint32_t *>(base + lay[8]); b.wide_count = ovf_count + 1;
    b.dp_list = reinterpret_cast<uint32_t *>(base + lay[9]); b.dp_count = ovf_count + 2; b.next_item = ovf_count + 3;
    b.wide_aux = reinterpret_cast<uint32_t *>(base + lay[10]);
    b.out = out; b.cigars = cigars; b.cigar_stride = cigar_stride;
    const unsigned task_blocks = (unsigned)((n + 127) / 128);
    const unsigned list_blocks = task_blocks < (unsigned)(4 * (sm_count > 0 ? sm_count : 148)) ? task_blocks : (unsigned)(4 * (sm_count > 0 ? sm_count : 148));
    // (1) gapless rectangles: one pass over the diagonal; (2) bands 1..3 and (3) bands 4..7 in registers, a thread per task;
    // (4) what they hand over (wider bands) one warp per task; (5) the serial kernel for the rest
    b.in_list = b.dp_list; b.in_count = b.dp_count;
    if (prm.use_pac) {
        { auto kern = sw_diag_kernel<true>; SALT_LAUNCH(kern, task_blocks, 128, 0, st, b); }
        { auto kern = sw_banded_narrow_kernel<true>; SALT_LAUNCH(kern, list_blocks, 128, 0, st, b); }
    } else {
        { auto kern = sw_diag_kernel<false>; SALT_LAUNCH(kern, task_blocks, 128, 0, st, b); }
        { auto kern = sw_banded_narrow_kernel<false>; SALT_LAUNCH(kern, list_blocks, 128, 0, st, b); }
    }
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    b.coop_list = reinterpret_cast<uint32_t *>(base + lay[11]); b.coop_aux = reinterpret_cast<uint32_t *>(base + lay[12]);
    b.coop_count = ovf_count + 4;
    b.in_list = b.wide_list; b.in_count = b.wide_count;
    if (prm.use_pac) { auto kern = sw_banded_medium_kernel<true>; SALT_LAUNCH(kern, list_blocks, 128, 0, st, b); }
    else { auto kern = sw_banded_medium_kernel<false>; SALT_LAUNCH(kern, list_blocks, 128, 0, st, b); }
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    b.in_list = b.coop_list; b.in_count = b.coop_count;
    {
        // one warp per handed-over task; scratch per warp: direction bytes of a band of COOP_MAXB (shared memory when they fit)
        const int rows8 = 8 * (((int)c.l_max + 7) / 8);
        const size_t cap = (size_t)(8 * (sm_count > 0 ? sm_count : 148));
        const size_t coop_blocks = (n + 3) / 4 < cap ? (n + 3) / 4 : cap;
        int smem_per_warp = (2 * COOP_MAXB + 1) * rows8;
        if (smem_per_warp > 11 * 1024) smem_per_warp = 11 * 1024;
        smem_per_warp = (smem_per_warp + 15) / 16 * 16;
        SALT_LAUNCH(sw_banded_coop_kernel, (unsigned)coop_blocks, 128, (size_t)4 * smem_per_warp, st, b, rows8, smem_per_warp);
    }
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    SALT_EV(5);

    if (ovf_dirs) {
        BandDev b2 = b;
        b2.dirs = ovf_dirs; b2.slot = (size_t)(2 * 512 + 1) * (size_t)(8 * (((int)c.l_max + 7) / 8));
        b2.in_list = b.ovf_list; b2.in_count = ovf_count;
        { auto kern = sw_banded_kernel<512, true>; SALT_LAUNCH(kern, (SSW_OVF_THREADS + 127) / 128, 128, 0, st, b2); }
        if ((e = cudaGetLastError()) != cudaSuccess) return e;
    }
    SALT_EV(6);
    if (launches) *launches += 9;
    return cudaSuccess;
#undef SALT_EV
}

}  // namespace salt
