// gm_machine.cuh -- the search kernels: tile staging, level-0 filters, and the
// search machine.
//
// gm_search_kernel<MODE, FULL, PF>, per tile (TILE consecutive nucleotides of
// the concatenated database, both strands):
//   1. one TMA bulk copy brings the packed tile + halo into shared memory; it is
//      expanded to one byte per nucleotide for the forward strand and for the
//      reverse complement (mk_rcmp, src/rnamot.c:193-216);
//   2. base bitsets (bit p of set x: base(p) == x) come from four ballots per 32
//      nucleotides; the reverse-complement strand's are their bit-reversed words,
//      and the "pair bitsets" of every distinct duplex table D (bit p of P[D][x]:
//      x pairs with the nucleotide at p) are unions of them, one word per lane;
//   3. the level-0 filter drops the starts that cannot begin a match:
//        PF 2  the sieve (sieve_word): the span-end test of the first helix for
//              32 starts per lane as word operations, with look-ahead bitsets for
//              the first / last helix of its interior and the literal prefilter
//              as further terms;
//        PF 0  per start: the 64-bit candidate mask of the first helix (wc_mask);
//        PF 1  the same plus the literal prefilter per start;
//   4. MODE 0 (fused): survivors go to a per-warp queue, lanes take them and run
//      the explicit-stack machine (gm_machine_body.inc) on the tile in shared
//      memory; MODE 1: survivors are appended to a global worklist that
//      gm_dfs_kernel consumes with per-lane windows (the worklist path, the
//      default where the filter is strong -- see DESIGN.md section 4).
// Every filter and mask only removes what match_wchlx would reject at its first
// tests (src/find_motif.c:1010-1079) or subtrees that cannot reach the hit sink;
// every survivor goes through the full test, so the enumeration is unchanged.
#pragma once

#include "gm_kernel.cuh"
#include "gm_tilebits.h"

namespace gm {

// Phases are grouped into classes (class = phase >> 3).  Each trip round the
// machine loop runs ONE class: the one most lanes of the warp are waiting in
// (majority vote).  Lanes of other classes sit the trip out, which lets lanes
// pile up in a class and then move through the phase cycle as a convoy instead
// of the warp executing every class with three or four lanes each.
enum {
	CL_IDLE = 0, CL_SPAN = 1, CL_WX = 2, CL_PK = 3, CL_TR = 4, CL_QU = 5
};
enum {
	PHX_IDLE = 0,
	PHX_ENTER = CL_SPAN * 8, PHX_SPAN, PHX_SS_RESUME, PHX_PH_RESUME, PHX_WC_RESUME,
	PHX_WX_BEGIN = CL_WX * 8, PHX_WX_RESUME, PHX_WX_FIRST, PHX_WX_EXT,
	PHX_PK_S5 = CL_PK * 8, PHX_PK_S3, PHX_PK_SD,
	PHX_TR_RESUME = CL_TR * 8, PHX_TR_S,
	PHX_QU_S1 = CL_QU * 8, PHX_QU_RESUME, PHX_QU_S2
};

#define PH_IDLE PHX_IDLE
#define PH_ENTER PHX_ENTER
#define PH_SPAN PHX_SPAN
#define PH_SS_RESUME PHX_SS_RESUME
#define PH_PH_RESUME PHX_PH_RESUME
#define PH_WC_RESUME PHX_WC_RESUME
#define PH_WX_BEGIN PHX_WX_BEGIN
#define PH_WX_RESUME PHX_WX_RESUME
#define PH_WX_FIRST PHX_WX_FIRST
#define PH_WX_EXT PHX_WX_EXT
#define PH_PK_S5 PHX_PK_S5
#define PH_PK_S3 PHX_PK_S3
#define PH_PK_SD PHX_PK_SD
#define PH_TR_RESUME PHX_TR_RESUME
#define PH_TR_S PHX_TR_S
#define PH_QU_S1 PHX_QU_S1
#define PH_QU_RESUME PHX_QU_RESUME
#define PH_QU_S2 PHX_QU_S2

// "return": the level is exhausted, its parent resumes (src/find_motif.c: every
// find_* returns to its caller's loop)
#define GM_RETURN()                                  \
	do {                                             \
		if (s == 0)                                  \
			ph = PH_IDLE;                            \
		else {                                       \
			s--;                                     \
			ph = hi16(L_FR(L, s, 1));                \
		}                                            \
	} while (0)

// frame word 1, low half: mpr (8 bits) | l_bpr << 8 | chk << 9
#define FR1_LO(mpr, lbpr, chk) (((mpr) & 0xff) | ((lbpr) << 8) | ((chk) << 9))

#define GM_QCAP 128 // per-warp queue of start items (power of two)
// Lanes are handed new starts in batches: a warp refills only once this many
// lanes are idle.  The lanes of a batch begin at search 0 together and move
// through the first levels in step (measured: 11.7 -> 12.7 G strand-nt/s on trna).
#define GM_REFILL_MIN 16

// per-warp tile bookkeeping (each warp owns its tile: buffers, mbarrier, queue)
struct WarpTile {
	uint64_t bar;
	int r_lo;          // first record intersecting the tile
	int one_rec;       // the tile lies inside a single record
};

__device__ __forceinline__ uint8_t expand_code(unsigned c)
{
	// low nibble IUPAC code, high nibble reference base code (rm_b2bc)
	unsigned b = c == 1 ? 0 : c == 2 ? 1 : c == 4 ? 2 : c == 8 ? 3 : 4;
	return (uint8_t)(c | (b << 4));
}
__device__ __forceinline__ uint8_t complement_byte(uint8_t v)
{
	// mk_rcmp, src/rnamot.c:200-208: a<->t, c<->g, everything else -> n
	unsigned b = v >> 4;
	if (b > 3)
		return (uint8_t)(15 | (4 << 4));
	unsigned nb = 3 - b;
	return (uint8_t)((1u << nb) | (nb << 4));
}

// 64 bits of a bitset starting at bit q (q >= 0; the set is padded by 4 words)
__device__ __forceinline__ uint64_t bits64(const uint32_t *set_, int q)
{
	const uint32_t *set = gm_sh(set_); // every bitset lives in shared memory
	const int w = q >> 5, sh = q & 31;
	const uint32_t a = set[w], b = set[w + 1], c = set[w + 2];
	const uint32_t lo = __funnelshift_r(a, b, sh);
	const uint32_t hi = __funnelshift_r(b, c, sh);
	return (uint64_t)lo | ((uint64_t)hi << 32);
}

struct PairBits {
	const uint32_t *base;  // [strand][dup][x][nwb]
	int nwb;               // words per bitset
	int n_dups;
};

// Span ends s3 in [lo, lo+n) (n <= 64) at which the helix whose 5' strand
// starts at window position z can have its first `req` pairs formed with at
// most `budget` mispairs -- a superset of the span ends match_wchlx accepts.
// bit j of the result <-> s3 = lo + j.
__device__ __forceinline__ uint64_t wc_mask_body(const PairBits &pb, const uint8_t *sq, int strand, int sqbase,
	int dupi, int flt, int z, int lo, int n)
{
	const uint64_t ones = n >= 64 ? ~0ull : ((1ull << n) - 1);
	const int req = flt & 0xff, budget = (flt >> 8) & 0xff, first_must = (flt >> 16) & 1;
	if (dupi < 0 || req == 0)
		return ones;
	const uint32_t *sets = pb.base + ((size_t)(strand * pb.n_dups + dupi) * 4) * pb.nwb;
	if (budget == 0 && first_must) {
		// the common case (no mispairs, paired ends): a plain AND chain
		uint64_t a = ones;
		for (int k = 0; k < req; k++) {
			const int x = bcode_of(sq[z + k]);
			if (x >= 4)
				return 0;
			a &= bits64(sets + x * pb.nwb, sqbase + lo - k);
		}
		return a;
	}
	uint64_t a0 = ones, a1 = ones, a2 = ones;
	for (int k = 0; k < req; k++) {
		const int x = bcode_of(sq[z + k]);
		uint64_t m = 0;
		if (x < 4)
			m = bits64(sets + x * pb.nwb, sqbase + lo - k);
		if (k == 0 && first_must) {
			a0 = a1 = a2 = m;
		} else if (k == 0 && budget == 0) {
			// ends without 5' pairing: an outermost mispair is tolerated even with
			// no mispair budget (src/find_motif.c:1014-1017 does not test mplim
			// there); with a budget it simply counts, as below
		} else {
			a2 = (a2 & m) | a1;
			a1 = (a1 & m) | a0;
			a0 &= m;
		}
	}
	return (budget == 0 ? a0 : budget == 1 ? a1 : a2) & ones;
}

// Out of line for the lite machine and the filters (their hot loops have to stay
// small: the kernels are instruction-cache sensitive); the full machine, whose
// pseudoknot loops call it per 3' end, inlines the body.
__device__ __noinline__ uint64_t wc_mask(const PairBits &pb, const uint8_t *sq, int strand, int sqbase,
	int dupi, int flt, int z, int lo, int n)
{
	return wc_mask_body(pb, sq, strand, sqbase, dupi, flt, z, lo, n);
}

// The level-0 sieve: the span-end test of wc_mask for 32 consecutive helix
// starts at once, one lane per 32-bit word of starts, word operations only.
// With M_d = { p : nucleotides p and p + d can pair } (built from the pair
// bitsets P[x] and the base bitsets I[x] = { p : base(p) = x }) the helix that
// starts at z and spans to z + D has its pair k at (z + k, z + D - k), i.e. bit
// z + k of M_{D-2k}.  Going through d upwards, M_d is computed once and ANDed
// (shifted by k) into the running products of the eight span offsets D = d + 2k
// it belongs to; the product of D = d is complete at that step.  Bit t of the
// result: some span offset in [Dlo, Dhi] lets the first `req` pairs of a helix
// starting at tile position 32 w + t form within the budget (0 or 1) -- what
// wc_mask(...) != 0 says for that start when the window is not clipped, and a
// superset of it when it is.
// `fin(d, f)` sees every completed span offset d >= Dlo with its word f and
// returns what of it counts (the identity, a look-ahead filter, ...).
template <int REQ, typename Fin>
__device__ __forceinline__ uint32_t sieve_word_r(const PairBits &pb, int strand, int dupi, int flt, int w, int Dlo, int Dhi, Fin fin)
{
	constexpr int NS_ = REQ > 0 ? REQ : 8; // accumulator slots
	const int req = REQ > 0 ? REQ : (flt & 0xff);
	const int budget = (flt >> 8) & 0xff, first_must = (flt >> 16) & 1;
	const int nwb = pb.nwb;
	const uint32_t *P = gm_sh(pb.base) + ((size_t)(strand * pb.n_dups + dupi) * 4) * nwb;
	const uint32_t *I = gm_sh(pb.base) + ((size_t)(strand * pb.n_dups) * 4) * nwb; // table 0: base bitsets
	uint32_t Bl[4], Bh[4];
#pragma unroll
	for (int x = 0; x < 4; x++) {
		Bl[x] = I[x * nwb + w];
		Bh[x] = I[x * nwb + w + 1];
	}
	uint32_t res = 0;
	const int Dmin = Dlo - 2 * (req - 1);
	for (int par = 0; par < 2; par++) {
		// a0[k] / a1[k]: products (no mispair / at most one) of the span offset that
		// is k steps from completion; the loop is kept rolled (rotating the
		// registers) so that it stays inside the L0 instruction cache
		uint32_t a0[NS_], a1[NS_];
#pragma unroll
		for (int u = 0; u < NS_; u++)
			a0[u] = a1[u] = ~0u;
		for (int d = Dmin + par; d <= Dhi; d += 2) {
			const int ww = min(w + (d >> 5), nwb - 3), sh = d & 31;
			uint32_t Ml = 0, Mh = 0;
#pragma unroll
			for (int x = 0; x < 4; x++) {
				const uint32_t c0 = P[x * nwb + ww], c1 = P[x * nwb + ww + 1], c2 = P[x * nwb + ww + 2];
				Ml |= Bl[x] & __funnelshift_r(c0, c1, sh);
				Mh |= Bh[x] & __funnelshift_r(c1, c2, sh);
			}
#pragma unroll
			for (int k = NS_ - 1; k >= 1; k--) {
				if (REQ > 0 || k < req) {
					const uint32_t t = __funnelshift_r(Ml, Mh, k);
					a1[k] = (a1[k] & t) | a0[k];
					a0[k] &= t;
				}
			}
			// the outermost pair closes span offset D = d (see wc_mask for the rules)
			uint32_t f;
			if (first_must)
				f = (budget ? a1[0] : a0[0]) & Ml;
			else if (budget == 0)
				f = a0[0];
			else
				f = (a1[0] & Ml) | a0[0];
			if (d >= Dlo)
				res |= fin(d, f);
#pragma unroll
			for (int u = 0; u < NS_ - 1; u++) {
				a0[u] = a0[u + 1];
				a1[u] = a1[u + 1];
			}
			a0[NS_ - 1] = a1[NS_ - 1] = ~0u;
		}
	}
	return res;
}

template <typename Fin>
__device__ __forceinline__ uint32_t sieve_word(const PairBits &pb, int strand, int dupi, int flt, int w, int Dlo, int Dhi, Fin fin)
{
	return sieve_word_r<0>(pb, strand, dupi, flt, w, Dlo, Dhi, fin);
}
// the main pass: the number of required pairs as a compile-time constant (no
// predicates, no dead accumulator slots in the hot loop)
template <typename Fin>
__device__ __forceinline__ uint32_t sieve_word_main(const PairBits &pb, int strand, int dupi, int flt, int w, int Dlo, int Dhi, Fin fin)
{
	switch (flt & 0xff) {
	case 3: return sieve_word_r<3>(pb, strand, dupi, flt, w, Dlo, Dhi, fin);
	case 4: return sieve_word_r<4>(pb, strand, dupi, flt, w, Dlo, Dhi, fin);
	case 5: return sieve_word_r<5>(pb, strand, dupi, flt, w, Dlo, Dhi, fin);
	case 6: return sieve_word_r<6>(pb, strand, dupi, flt, w, Dlo, Dhi, fin);
	case 7: return sieve_word_r<7>(pb, strand, dupi, flt, w, Dlo, Dhi, fin);
	}
	return sieve_word_r<0>(pb, strand, dupi, flt, w, Dlo, Dhi, fin);
}

// 32 bits of a bitset from bit q on (q >= 0)
__device__ __forceinline__ uint32_t bits32(const uint32_t *set_, int q)
{
	const uint32_t *set = gm_sh(set_);
	return __funnelshift_r(set[q >> 5], set[(q >> 5) + 1], q & 31);
}

// The mirror image of wc_mask: the helix's 3' END e is known and its 5' start is
// not.  5' starts zt in [zlo, zlo+n) (n <= 64) at which the first `req` pairs
// (zt+k, e-k) can form within the budget; bit j <-> zt = zlo + j.  `dupi_t`
// indexes the bitsets of the TRANSPOSED duplex table (bit p of set y = the
// nucleotide at p, as 5' base, pairs with 3' base y).
__device__ __noinline__ uint64_t wc_mask_rev(const PairBits &pb, const uint8_t *sq, int strand, int sqbase,
	int dupi_t, int flt, int e, int zlo, int n)
{
	const uint64_t ones = n >= 64 ? ~0ull : ((1ull << n) - 1);
	const int req = flt & 0xff, budget = (flt >> 8) & 0xff, first_must = (flt >> 16) & 1;
	if (dupi_t < 0 || req == 0)
		return ones;
	const uint32_t *sets = pb.base + ((size_t)(strand * pb.n_dups + dupi_t) * 4) * pb.nwb;
	uint64_t a0 = ones, a1 = ones, a2 = ones;
	for (int k = 0; k < req; k++) {
		const int y = bcode_of(sq[e - k]);
		uint64_t m = 0;
		if (y < 4)
			m = bits64(sets + y * pb.nwb, sqbase + zlo + k);
		if (k == 0 && first_must) {
			a0 = a1 = a2 = m;
		} else if (k == 0 && budget == 0) {
			// see wc_mask: an outermost mispair is tolerated without a budget
		} else {
			a2 = (a2 & m) | a1;
			a1 = (a1 & m) | a0;
			a0 &= m;
		}
	}
	return (budget == 0 ? a0 : budget == 1 ? a1 : a2) & ones;
}

// Tail look-ahead (DevSearch::lk_t): with helix S chosen as (s5, s3, hl), can the
// helix T whose 3' strand ends lk_off nucleotides before S's 3' strand form at all?
// T ends at e = s3 - hl - lk_off and begins somewhere in [e - dhi, e - dlo], not
// before S's 5' strand ends.  Pure pruning: a false answer means no
// assignment of the interior reaches the hit sink.
__device__ __forceinline__ bool tail_feasible(const PairBits &pb, const uint8_t *sq, int strand, int sqbase,
	const DevSearch &S, const DevSearch &T, int s5, int s3, int hl)
{
	const int e = s3 - hl - S.lk_off;
	int zhi = e - T.dlo;
	const int zlo = max(e - T.dhi, s5 + hl);
	for (; zhi >= zlo; zhi -= 64) {
		const int l0 = max(zlo, zhi - 63);
		if (wc_mask_rev(pb, sq, strand, sqbase, T.dupi_t, T.flt, e, l0, zhi - l0 + 1) != 0)
			return true;
	}
	return false;
}

// Forward look-ahead: can helix T, whose 5' strand starts at zt, have any span end
// in [zt + dlo, min(zt + dhi, bound)] at which its first pairs form?
#define GM_LOOK_FWD(any, T, zt, bound)                                   \
	do {                                                                 \
		int f_ = min((bound), (zt) + (T).dhi);                           \
		const int l_ = (zt) + (T).dlo;                                   \
		(any) = false;                                                   \
		for (; f_ >= l_ && !(any); f_ -= 64) {                           \
			const int l0_ = max(l_, f_ - 63);                            \
			(any) = GM_MASK((T), (zt), l0_, f_ - l0_ + 1) != 0;          \
		}                                                                \
	} while (0)

// Probes (DevSearch::probe): fixed-length single strands with seq= whose place follows
// from the helix (s5, s3, hl) just chosen.  wend = last position of the window.
__device__ __noinline__ bool probes_ok(const Lane &L, const DevSearch &S, int s5, int s3, int hl, int wend)
{
	for (int i = 0; i < S.n_probe; i++) {
		const unsigned pr = S.probe[i];
		const int anchor = pr & 3, off = (pr >> 2) & 1023, len = (pr >> 12) & 255;
		const int rxi = (pr >> 20) & 31, mm = (pr >> 25) & 15;
		const int pos = anchor == 0 ? s5 + hl + off : anchor == 1 ? s3 + 1 + off : s3 - hl + 1 - off - len;
		if (pos < 0 || pos + len - 1 > wend)
			return false;
		const DevRegex &rx = PV.regex[rxi];
		int n_mm;
		if (mm > 0 ? !rx_match_mm(rx, L.sq + pos, len, mm, &n_mm) : !rx_match(rx, L.sq + pos, len))
			return false;
	}
	return true;
}

// MODE 0: fused -- prefilter and machine in one kernel (works for every plan).
// MODE 1: prefilter only -- survivors are appended to a global worklist that
//         gm_dfs_kernel consumes (the split path).
//
// Every warp works on its own: it pulls tiles from the global counter, stages
// them into one of ITS TWO tile buffers with its own TMA bulk copy + mbarrier,
// and keeps its lanes fed from a private queue of prefilter survivors.  When a
// tile's starts are used up the warp stages the next tile into the other
// buffer while lanes that are still enumerating on the old one carry on, so
// neither a block-wide barrier nor the end of a tile ever idles the lanes
// (a buffer is recycled only once no lane works in it).
// FULL = false compiles the machine for plans made of single strands and proper
// helices only (most descriptors): the pseudoknot / parallel helix / triplex /
// quadruplex code is left out, which keeps the hot loop inside the
// instruction cache.
// PF selects the level-0 prefilter: 0 = per-start candidate mask, 1 = the same
// plus the literal prefilter (plans with a gm_plan_t::literal), 2 = the
// word-parallel sieve (sieve_word).  A template parameter so that each plan
// runs only the code it needs (the kernel is instruction-cache bound).
template <int MODE, bool FULL, int PF>
__device__ __forceinline__ void gm_search_body(const ScanArgs &A)
{
	constexpr bool SIEVE = PF >= 2; // word-parallel level-0 sieve instead of the per-start prefilter (3: two-stage)
	// literal prefilter: per start (PF == 1) or as one more term of the sieve
	const bool LIT = PF == 1 || (SIEVE && A.par.lit_present != 0);
	extern __shared__ __align__(16) uint8_t smem_raw[];
	const int tid = threadIdx.x, nt = blockDim.x;
	const int lane = tid & 31, warp = tid >> 5;
	const int NS = A.par.n_searches, ND = A.par.n_descr;
	const int W = A.par.w_winsize, H = A.par.halo, TILE = A.par.tile;
	const int Lbytes = (TILE + 2 * H + 31) & ~31;          // nucleotides staged per tile (whole bitset words)
	const int stage_bytes = ((Lbytes >> 1) + 32 + 15) & ~15; // packed staging (+ alignment slack)
	const int nwb = ((Lbytes + 31) >> 5) + 4;
	const int n_dups = A.par.n_dups;
	const int NBUF = MODE == 0 ? 2 : 1;

	// carve shared memory (mirrors smem_need() on the host): the staged plan shared
	// by the block, then one private region per warp, then the lane state
	const int nwarps = nt >> 5;
	const size_t pb_bytes = (((size_t)2 * n_dups * 4 * nwb * 4) + 15) & ~(size_t)15;
	const size_t lit_bytes = LIT ? (((size_t)2 * nwb * 4) + 15) & ~(size_t)15 : 0;
	const size_t buf_bytes = 2 * (size_t)Lbytes + pb_bytes + (GM_REC_CACHE + 2) * 8 + lit_bytes;
	const bool CHAIN = SIEVE && A.par.chain > 0;
	const size_t warp_bytes = 16 + (size_t)stage_bytes + NBUF * buf_bytes + GM_QCAP * 2 + (SIEVE ? ((6 * (size_t)nwb * 4 + 15) & ~(size_t)15) : 0) +
		(CHAIN ? ((6 * (size_t)nwb * 4 + 15) & ~(size_t)15) : 0);
	StagedPlan sp;
	uint8_t *p = stage_plan(smem_raw, A, tid, nt, sp);
	DevSearch *sm_ds = sp.ds;
	uint64_t *sm_litB = reinterpret_cast<uint64_t *>(p);       p += LIT ? 16 * 8 : 0; // literal prefilter class masks
	uint8_t *wp = p + (size_t)warp * warp_bytes;               p += (size_t)nwarps * warp_bytes;
	uint32_t *sm_state = reinterpret_cast<uint32_t *>(p);
	WarpTile *sm = reinterpret_cast<WarpTile *>(wp);           wp += 16;
	uint8_t *sm_stage = wp;                                    wp += stage_bytes;
	uint8_t *bufs = wp;                                        wp += NBUF * buf_bytes;
	uint16_t *myq = reinterpret_cast<uint16_t *>(wp);
	// look-ahead bitsets of the sieve (DevParams::pf_deep), per strand: sv_E = ends
	// at which the last helix of the first helix's interior can form, sv_K =
	// starts at which its first helix can
	uint32_t *sv_E = reinterpret_cast<uint32_t *>(myq + GM_QCAP);
	uint32_t *sv_K = sv_E + 2 * nwb;
	uint32_t *sv_K2 = sv_K + 2 * nwb; // starts at which the helix FOLLOWING that first helix can form (pf_deep == 2)
	// composition chain (DevParams::chain): two feasibility bitsets per strand
	// (ping-pong) and the bitset of positions whose base the current element allows
	uint32_t *ch_F = reinterpret_cast<uint32_t *>(reinterpret_cast<uint8_t *>(sv_E) + (SIEVE ? ((6 * (size_t)nwb * 4 + 15) & ~(size_t)15) : 0));
	uint32_t *ch_OK = ch_F + 4 * nwb;
	const uint32_t *ch_F0 = ch_F; // the finished chain: bit p <=> the whole descriptor can be laid out from p
	const bool deep = SIEVE && A.par.pf_deep != 0;
	constexpr bool TWO = MODE == 1 && PF == 3; // two-stage sieve (worklist path only): its own instantiation,
	                                           // without the word-parallel main pass and the end bitsets
	// per-start stage behind the sieve words (accept2): always in two-stage mode, and when
	// the first helix has probes (seq= of single strands at a place the helix fixes)
	const bool ST2 = TWO || (MODE == 1 && SIEVE && A.par.sv_helix != 0 && A.par.pf_search >= 0 &&
		A.ds[max(A.par.pf_search, 0)].n_probe > 0);
	const int sv_nws = ((TILE - 1) >> 5) + 2;               // sieve words per strand
	const int sv_npass = (A.strands * sv_nws + 31) >> 5;

	if (LIT && tid < 16)
		sm_litB[tid] = A.plan->regex[A.par.lit_rx].B[tid];
	if (lane == 0)
		mbar_init(&sm->bar, 1);

	Lane L;
	L.P = sp.pv;
	L.st = sm_state + tid;
	L.nt = nt;
	L.ds = sm_ds;
	L.ps = sp.ps;
	L.NS = NS;
	L.ND = ND;
	L.el_base = NS + A.par.frame_words;
	L.sq = bufs;
	L.szero = L.slen = L.comp = 0;
	L.rec = 0;
	L.seq = 0;
	// never-marked elements read as UNDEF; counters start at UNDEF like
	// SE_init leaves them (src/compile.c:570-571)
	if (MODE == 0) {
		for (int d = 0; FULL && d < ND; d++) {
			unmark(L, d);
			set_cnt(L, d, GM_UNDEF, GM_UNDEF);
		}
	}
	__syncthreads(); // the only block-wide barrier: plan tables are staged

	uint32_t parity = 0;
	unsigned long long my_starts = 0;
	unsigned my_entries = 0; // starts handed to the machine (debug statistics)

	// current tile (warp-uniform)
	int cur = NBUF - 1;          // buffer of the current tile (first load flips it to 0)
	bool have_tile = false, no_more_tiles = false;
	int64_t gA = 0, gB = 0, lo = 0;
	uint8_t *sm_fwd = bufs, *sm_rc = bufs;
	int64_t *sm_rec = reinterpret_cast<int64_t *>(bufs);
	PairBits pb;
	pb.base = reinterpret_cast<uint32_t *>(bufs);
	pb.nwb = nwb;
	pb.n_dups = n_dups;
	uint32_t *sm_lit = reinterpret_cast<uint32_t *>(bufs); // literal-occurrence bitsets [strand][nwb]
	bool one_rec = false;
	int r_lo = 0;
	int64_t rec0_off = 0;
	int rec0_len = 0;
	int work_next = 0;
	const int n_work = A.strands * TILE;
	const int refill_min = A.par.refill_min;

	// per lane: where the start it is enumerating lives
	int sqbase = 0, strand = 0, mybuf = 0;
	PairBits mypb = pb;

	// stage tile t into buffer b
	auto load_tile = [&](int b, int64_t t) {
		uint8_t *bp = bufs + (size_t)b * buf_bytes;
		sm_fwd = bp;
		sm_rc = bp + Lbytes;
		pb.base = reinterpret_cast<uint32_t *>(bp + 2 * (size_t)Lbytes);
		sm_rec = reinterpret_cast<int64_t *>(bp + 2 * (size_t)Lbytes + pb_bytes);
		sm_lit = reinterpret_cast<uint32_t *>(bp + 2 * (size_t)Lbytes + pb_bytes + (GM_REC_CACHE + 2) * 8);
		gA = A.g_begin + t * (int64_t)TILE;
		gB = min(gA + (int64_t)TILE, A.g_end);
		lo = gA - H;
		// packed bytes [bs, bs + nbytes) cover nucleotides [lo_c, hi_c)
		const int64_t lo_c = max(lo, (int64_t)0);
		const int64_t hi_c = min(lo + Lbytes, A.total_nt);
		const int64_t bs = (lo_c >> 1) & ~(int64_t)15;
		const uint32_t nbytes = (uint32_t)((((hi_c + 1) >> 1) - bs + 15) & ~(int64_t)15);
		if (lane == 0) {
			mbar_expect_tx(&sm->bar, nbytes);
			tma_bulk_g2s(sm_stage, A.packed + bs, nbytes, &sm->bar);
			// the last record starting at or before gA
			int a = 0, bb = A.n_rec; // rec_off[a] <= gA < rec_off[bb]
			while (bb - a > 1) {
				int m = (a + bb) >> 1;
				if (A.rec_off[m] <= gA)
					a = m;
				else
					bb = m;
			}
			sm->r_lo = a;
			sm->one_rec = A.rec_off[a + 1] >= gB;
		}
		__syncwarp();
		r_lo = sm->r_lo;
		one_rec = sm->one_rec != 0;
		// cache the offsets of up to GM_REC_CACHE records from r_lo on
		for (int i = lane; i <= GM_REC_CACHE; i += 32) {
			int r = r_lo + i;
			sm_rec[i] = r <= A.n_rec ? A.rec_off[r] : (int64_t)1 << 62;
		}
		mbar_wait(&sm->bar, parity);
		parity ^= 1;
		// expand packed nibbles to one byte per nucleotide, both strands, and build
		// the forward base bitsets (set 0 = the identity table: bit p of set x says
		// base(p) == x) with four ballots per 32 nucleotides
		uint32_t *pbw = const_cast<uint32_t *>(pb.base);
		const int nw = Lbytes >> 5;
		{
			// Eight nucleotides per lane and trip (expand8, gm_tilebits.h: SIMD within a
			// register, no branches): one funnel shift brings the lane's eight codes in
			// line whatever the parity of the tile's first nucleotide; the four lanes of
			// a 32-nucleotide word then exchange the bytes of their base-bitset
			// contributions (two shuffles + byte permutes), lane 4k + b ending up with
			// base b's word.
			const uint32_t *stw = reinterpret_cast<const uint32_t *>(sm_stage);
			const int n_stw = stage_bytes >> 2;
			const int base_q = (int)(lo - 2 * bs); // nibble of tile position 0 in the staging buffer (< 0 before the database)
			const int sh = (base_q & 7) * 4;
			const bool edge = lo < 0 || lo + Lbytes > A.total_nt; // positions outside the database read as code 0
			const unsigned sel1 = (lane & 1) ? 0x3715u : 0x6240u, sel2 = (lane & 2) ? 0x3276u : 0x5410u;
			for (int i0 = lane * 8; i0 < ((Lbytes + 255) & ~255); i0 += 256) {
				const int wq = (base_q + i0) >> 3;
				const uint32_t w0 = stw[min(max(wq, 0), n_stw - 1)], w1 = stw[min(max(wq + 1, 0), n_stw - 1)];
				uint32_t x = __funnelshift_r(w0, w1, sh);
				if (edge) {
					const int64_t g0 = lo + i0;
					const int klo = (int)min((int64_t)8, max((int64_t)0, -g0));
					const int khi = (int)min((int64_t)8, max((int64_t)0, A.total_nt - g0));
					const uint32_t m_hi = khi >= 8 ? ~0u : ((1u << (4 * khi)) - 1u);
					const uint32_t m_lo = klo >= 8 ? 0u : (~0u << (4 * klo));
					x &= m_hi & m_lo;
				}
				uint32_t f0, f1, r0, r1, bits;
				expand8(x, f0, f1, r0, r1, bits);
				uint32_t o = __shfl_xor_sync(0xffffffffu, bits, 1);
				bits = __byte_perm(bits, o, sel1);
				o = __shfl_xor_sync(0xffffffffu, bits, 2);
				bits = __byte_perm(bits, o, sel2);
				if (i0 < Lbytes) {
					*reinterpret_cast<uint2 *>(sm_fwd + i0) = make_uint2(f0, f1);
					*reinterpret_cast<uint2 *>(sm_rc + Lbytes - 8 - i0) = make_uint2(r0, r1);
					pbw[(size_t)(lane & 3) * nwb + (i0 >> 5)] = bits;
				}
			}
		}
		if (lane < 4 * (nwb - nw))
			pbw[(size_t)(lane / (nwb - nw)) * nwb + nw + lane % (nwb - nw)] = 0; // padding words
		__syncwarp();
		// reverse-complement base bitsets: position i of that strand holds the
		// complement of forward position Lbytes-1-i, so word w of set x is the
		// bit-reversed word nw-1-w of forward set 3-x
		for (int x = 0; x < 4; x++)
			for (int w = lane; w < nwb; w += 32)
				pbw[((size_t)(n_dups * 4) + x) * nwb + w] = w < nw ? __brev(pbw[(size_t)(3 - x) * nwb + nw - 1 - w]) : 0u;
		__syncwarp();
		// pair bitsets of every other table: set x = union of the base bitsets of
		// the bases x pairs with, one word per lane
		for (int sd = 0; sd < 2 * n_dups; sd++) {
			const int st = sd >= n_dups, dd = st ? sd - n_dups : sd;
			if (dd == 0)
				continue;
			const unsigned dup = A.par.dups[dd];
			const uint32_t *bs_ = pbw + (size_t)(st * n_dups) * 4 * nwb;
			uint32_t *out = pbw + (size_t)(st * n_dups + dd) * 4 * nwb;
			for (int x = 0; x < 4; x++) {
				const unsigned row = dup >> (x * 5);
				for (int w = lane; w < nwb; w += 32) {
					uint32_t acc = 0;
#pragma unroll
					for (int y = 0; y < 4; y++)
						if ((row >> y) & 1u)
							acc |= bs_[(size_t)y * nwb + w];
					out[(size_t)x * nwb + w] = acc;
				}
			}
		}
		if (LIT) {
			// literal prefilter (adjust_szero, src/find_motif.c:209-243): bit i of a
			// strand's set = the best literal can occur at tile position i within its
			// mismatch allowance (mm_advance on fixed-length items, src/mm_regexp.c:369-469).
			// Word arithmetic over the base bitsets, one word per lane: pattern position
			// k accepts base x if its class mask says so, and any nucleotide that is not
			// plain a/c/g/t counts as accepted -- a superset of the true occurrences
			// (survivors take the exact tests in the machine).
			const int len = A.par.lit_len, l_mm = A.par.lit_mm;
			const uint64_t dot = PV.regex[A.par.lit_rx].dot;
			const uint64_t sel0 = sm_litB[1], sel1 = sm_litB[2], sel2 = sm_litB[4], sel3 = sm_litB[8];
			for (int idx = lane; idx < 2 * nwb; idx += 32) {
				const int st = idx >= nwb, w = idx - st * nwb;
				const uint32_t *Bs = pbw + (size_t)(st * n_dups) * 4 * nwb;
				uint32_t a0 = ~0u, a1 = ~0u, a2 = ~0u;
				for (int k = 0; k < len; k++) {
					if ((dot >> k) & 1)
						continue;
					const int q = min((w << 5) + k, Lbytes);
					const uint32_t b0 = bits32(Bs, q), b1 = bits32(Bs + nwb, q), b2 = bits32(Bs + 2 * nwb, q),
						b3 = bits32(Bs + 3 * nwb, q);
					const uint32_t m = ~(b0 | b1 | b2 | b3) | (((sel0 >> k) & 1) ? b0 : 0u) | (((sel1 >> k) & 1) ? b1 : 0u) |
						(((sel2 >> k) & 1) ? b2 : 0u) | (((sel3 >> k) & 1) ? b3 : 0u);
					a2 = (a2 & m) | a1;
					a1 = (a1 & m) | a0;
					a0 &= m;
				}
				sm_lit[idx] = l_mm == 0 ? a0 : l_mm == 1 ? a1 : l_mm == 2 ? a2 : ~0u;
			}
		}
		__syncwarp();
		rec0_off = sm_rec[0];
		rec0_len = (int)(sm_rec[1] - sm_rec[0]);
		work_next = 0;
		cur = b;
		have_tile = true;
	};

	// next tile index for this warp, or -1
	auto next_tile = [&]() -> int64_t {
		unsigned long long t_ = 0;
		if (lane == 0)
			t_ = atomicAdd(A.tile_counter, 1ull);
		const int64_t t = (int64_t)__shfl_sync(0xffffffffu, t_, 0);
		return t < A.n_tiles ? t : -1;
	};

	// locate start item q of the current tile: false if it is not a start of this scan
	auto locate = [&](int q, int &comp, int &idx, uint32_t &rec, int &slen, int &szero) -> bool {
		if (q >= n_work)
			return false;
		comp = q >= TILE;
		const int64_t g = gA + (comp ? q - TILE : q);
		if (g >= gB)
			return false;
		int64_t off;
		if (one_rec) {
			off = rec0_off;
			slen = rec0_len;
			rec = (uint32_t)r_lo;
		} else if (g < sm_rec[GM_REC_CACHE]) {
			int a = 0, b = GM_REC_CACHE;
			while (b - a > 1) {
				int m = (a + b) >> 1;
				if (sm_rec[m] <= g)
					a = m;
				else
					b = m;
			}
			off = sm_rec[a];
			slen = (int)(sm_rec[a + 1] - off);
			rec = (uint32_t)(a + r_lo);
		} else {
			int a = r_lo, b = A.n_rec;
			while (b - a > 1) {
				int m = (a + b) >> 1;
				if (A.rec_off[m] <= g)
					a = m;
				else
					b = m;
			}
			off = A.rec_off[a];
			slen = (int)(A.rec_off[a + 1] - off);
			rec = (uint32_t)a;
		}
		const int pos = (int)(g - off);
		szero = comp ? slen - 1 - pos : pos;
		idx = (int)(g - lo);
		// RM_find_motif searches szero in [0, slen - rm_dminlen], src/find_motif.c:184-205
		return slen - szero >= A.par.dminlen;
	};

	// level-0 prefilter of start item q.  v0/have_v0: the candidate mask of
	// search 0's span ends when they all fit one 64-wide chunk.
	auto prefilter = [&](int q, uint64_t &v0, int &have_v0) -> bool {
		int comp, idx, slen, szero;
		uint32_t rec;
		v0 = 0;
		have_v0 = 0;
		bool pass = locate(q, comp, idx, rec, slen, szero);
		if (!pass)
			return false;
		my_starts++;
		const DevSearch &S0 = sm_ds[0];
		const uint8_t *sq = comp ? sm_rc + (Lbytes - 1 - idx) : sm_fwd + idx;
		const int base = comp ? Lbytes - 1 - idx : idx;
		const int dl = min(W, slen - szero) - 1;
		if (LIT) {
			// the literal must begin lmin..lmax nucleotides after the start and end
			// inside the window
			const int l = A.par.lit_lmin;
			int h = min(A.par.lit_lmax, dl + 1 - A.par.lit_len);
			bool any = false;
			const uint32_t *set = sm_lit + comp * nwb;
			for (; h >= l && !any; h -= 64) {
				const int l0 = max(l, h - 63), n = h - l0 + 1;
				const uint64_t ones = n >= 64 ? ~0ull : ((1ull << n) - 1);
				any = (bits64(set, base + l0) & ones) != 0;
			}
			if (!any)
				return false;
		}
		if (A.par.pf_search >= 0) {
			// any span end at all for the first helix of the descriptor?  (It is
			// search 0, or follows fixed-length single strands, so its 5' start
			// pz is known.)
			const DevSearch &SP = sm_ds[A.par.pf_search];
			const int pz = A.par.pf_z;
			const int fsd = min(dl, pz + SP.dhi), lsd = pz + SP.dlo;
			bool any = false;
			for (int hi = fsd; hi >= lsd && !any; hi -= 64) {
				const int l0 = max(lsd, hi - 63);
				const uint64_t v = wc_mask(pb, sq, comp, base, SP.dupi, SP.flt, pz, l0, hi - l0 + 1);
				any = v != 0;
				if (hi == fsd && l0 == lsd && SP.kind != K_PK && A.par.pf_search == 0) {
					v0 = v;
					have_v0 = 1;
				}
			}
			pass = any;
		}
		if (pass && S0.rx5 >= 0 && S0.mm5 == 0 && !PV.regex[S0.rx5].eol) {
			// a seq= without '$' that cannot match the longest
			// placement cannot match a shorter one
			pass = rx_match(PV.regex[S0.rx5], sq, min(S0.maxlen, dl + 1)) != 0;
		}
		return pass;
	};

	// what is left of the prefilter for a start the sieve let through: is it a
	// start of this scan at all, and can an anchored seq= of search 0 match
	auto accept = [&](int q) -> bool {
		int comp, idx, slen, szero;
		uint32_t rec;
		if (!locate(q, comp, idx, rec, slen, szero))
			return false;
		const DevSearch &S0 = sm_ds[0];
		if (S0.rx5 >= 0 && S0.mm5 == 0 && !PV.regex[S0.rx5].eol) {
			const uint8_t *sq = comp ? sm_rc + (Lbytes - 1 - idx) : sm_fwd + idx;
			const int dl = min(W, slen - szero) - 1;
			return rx_match(PV.regex[S0.rx5], sq, min(S0.maxlen, dl + 1)) != 0;
		}
		return true;
	};

	// Stage 2 of the two-stage sieve (DevParams::sv_two), per start: the first helix's
	// span-end test (wc_mask) and, for every span end and helix length that passes it,
	// the look-ahead the word-parallel main pass applies per span offset -- the first
	// interior helix can start behind the 5' strand (a bit of sv_K) and the last one
	// can end before the 3' strand (tail_feasible).
	auto accept2 = [&](int q) -> bool {
		int comp, idx, slen, szero;
		uint32_t rec;
		if (!locate(q, comp, idx, rec, slen, szero))
			return false;
		const DevSearch &S0 = sm_ds[0];
		const DevSearch &SP = sm_ds[A.par.pf_search];
		const uint8_t *sq = comp ? sm_rc + (Lbytes - 1 - idx) : sm_fwd + idx;
		const int base = comp ? Lbytes - 1 - idx : idx;
		const int dl = min(W, slen - szero) - 1;
		if (S0.rx5 >= 0 && S0.mm5 == 0 && !PV.regex[S0.rx5].eol &&
		    !rx_match(PV.regex[S0.rx5], sq, min(S0.maxlen, dl + 1)))
			return false;
		const int pz = A.par.pf_z;
		const int lsd = pz + SP.dlo;
		const uint32_t *K = sv_K + comp * nwb;
		const bool use_k = deep && SP.kid_t >= 0;
		for (int hi = min(dl, pz + SP.dhi); hi >= lsd; hi -= 64) {
			const int l0 = max(lsd, hi - 63);
			uint64_t v = wc_mask(pb, sq, comp, base, SP.dupi, SP.flt, pz, l0, hi - l0 + 1);
			while (v) {
				const int s3 = l0 + __ffsll((long long)v) - 1;
				v &= v - 1;
				for (int hl = SP.minlen; hl <= SP.maxlen; hl++) {
					const int kp = base + pz + hl + SP.kid_off;
					if (use_k && !((K[kp >> 5] >> (kp & 31)) & 1u))
						continue;
					if (SP.lk_t >= 0 && !tail_feasible(pb, sq, comp, base, SP, sm_ds[SP.lk_t], pz, s3, hl))
						continue;
					if (SP.n_probe) {
						Lane Lq = L;
						Lq.sq = sq;
						if (!probes_ok(Lq, SP, pz, s3, hl, dl))
							continue;
					}
					return true;
				}
			}
		}
		return false;
	};

	// ---- composition chain (a sieve term) ----
	// Going through the descriptor's elements from the last to the first, F(p) = "this
	// element and everything behind it can be laid out contiguously from p, every
	// helix strand made of bases its pair table can pair at all (up to the mispair
	// budget, end positions as `ends` demands)".  Single strands and strands any base
	// can sit in only dilate F by their length range.  Equal strand lengths and the
	// pairing itself are left to the machine: a superset of the starts with a
	// candidate, computed for 32 starts per word.  One word per lane.
	auto chain_build = [&]() {
		const int nw = Lbytes >> 5;
		uint32_t *Fa = ch_F, *Fb = ch_F + 2 * nwb;
		for (int i = lane; i < 2 * nwb; i += 32)
			Fa[i] = ~0u; // behind the last element anything goes
		__syncwarp();
		const uint32_t *pbw = pb.base;
		for (int c = 0; c < A.par.chain; c++) {
			const unsigned w0 = A.par.chain_w0[c], w1 = A.par.chain_w1[c];
			const int mn = w0 & 4095, mx = (w0 >> 12) & 4095, cm = (w0 >> 24) & 15;
			const bool both = (w0 >> 28) & 1, con = (w0 >> 29) & 1;
			if (con) {
				for (int i = lane; i < 2 * nwb; i += 32) {
					const int st = i >= nwb, w = i - st * nwb;
					const uint32_t *Bs = pbw + (size_t)(st * n_dups) * 4 * nwb;
					ch_OK[i] = ((cm & 1) ? Bs[w] : 0u) | ((cm & 2) ? Bs[nwb + w] : 0u) | ((cm & 4) ? Bs[2 * nwb + w] : 0u) |
						((cm & 8) ? Bs[3 * nwb + w] : 0u);
				}
				__syncwarp();
			}
			for (int i = lane; i < 2 * nwb; i += 32) {
				const int st = i >= nwb, w = i - st * nwb;
				uint32_t out = ~0u; // beyond the staged range: unknown, so feasible
				if (w < nw && st < A.strands) {
					const uint32_t *Fn = Fa + st * nwb;
					const int q0 = w << 5;
					out = 0;
					if (!con) {
						for (int len = mn; len <= mx && out != ~0u; len++)
							out |= bits32(Fn, min(q0 + len, Lbytes));
					} else {
						const uint32_t *OK = ch_OK + st * nwb;
						uint32_t a0 = ~0u, a1 = ~0u, a2 = ~0u; // at most 0 / 1 / 2 positions so far whose base is not allowed
						const uint32_t ok0 = OK[w];
						for (int k = 0; k < mx; k++) {
							const uint32_t m = bits32(OK, min(q0 + k, Lbytes));
							a2 = (a2 & m) | a1;
							a1 = (a1 & m) | a0;
							a0 &= m;
							const int hl = k + 1;
							if (hl >= mn) {
								const int b = (w1 >> (2 * (hl - mn))) & 3;
								uint32_t okh = b == 0 ? a0 : b == 1 ? a1 : b == 2 ? a2 : ~0u;
								if (both && b != 3)
									okh &= ok0 & m;
								out |= okh & bits32(Fn, min(q0 + hl, Lbytes));
							}
						}
					}
				}
				Fb[i] = out;
			}
			__syncwarp();
			uint32_t *t_ = Fa;
			Fa = Fb;
			Fb = t_;
		}
		ch_F0 = Fa;
	};

	// ---- the sieve (PF == 2), shared by the fused and the prefilter-only kernel ----
	// look-ahead bitsets for the current tile, one word per lane: over the range
	// of positions the main pass can ask about
	auto sieve_aux = [&]() {
		const DevSearch &SP = sm_ds[A.par.pf_search];
		const int nw = Lbytes >> 5;
		for (int i = lane; i < 2 * nwb; i += 32) {
			sv_E[i] = 0;
			sv_K[i] = 0;
			sv_K2[i] = 0;
		}
		__syncwarp();
		const int nv = (int)(gB - gA);
		for (int st = 0; st < A.strands; st++) {
			const int zlo = (st ? Lbytes - H - nv : H) + A.par.pf_z, zhi = zlo + nv - 1; // helix starts of this tile
			if (SP.lk_t >= 0 && !TWO) {
				// ends asked about: zb + d - hl - lk_off; the helices that end there start
				// up to maxglen - 1 earlier
				const DevSearch &T = sm_ds[SP.lk_t];
				const int lo_ = max(zlo + SP.dlo - SP.maxlen - SP.lk_off - T.dhi, 0);
				const int hi_ = min(zhi + SP.dhi - SP.minlen - SP.lk_off - T.dlo, Lbytes - 1);
				uint32_t *E = sv_E + st * nwb;
				for (int w = (lo_ >> 5) + lane; w <= (hi_ >> 5) && w < nw; w += 32)
					sieve_word_main(pb, st, T.dupi, T.flt, w, T.dlo, T.dhi,
						[&](int d, uint32_t f) -> uint32_t {
							// the helix that starts at bit t ends at t + d
							const int q = (w << 5) + d, sh = q & 31;
							if (f != 0) {
								if ((q >> 5) < nwb)
									atomicOr(E + (q >> 5), f << sh);
								if (sh && (q >> 5) + 1 < nwb)
									atomicOr(E + (q >> 5) + 1, f >> (32 - sh));
							}
							return 0u;
						});
			}
			if (SP.kid_t >= 0) {
				const DevSearch &T = sm_ds[SP.kid_t];
				const int lo_ = zlo + SP.minlen + SP.kid_off;
				const int hi_ = min(zhi + SP.maxlen + SP.kid_off + 31, Lbytes - 1);
				const bool has_k2 = A.par.pf_deep == 2;
				const uint32_t *K2 = sv_K2 + st * nwb;
				if (has_k2) {
					// the helix after the first interior helix starts sib_off + 1 behind its
					// group: sieve its starts first, over everything the next loop asks about
					const DevSearch &T2 = sm_ds[T.sib_t];
					const int lo2 = (lo_ & ~31) + T.dlo + 1 + T.sib_off;
					const int hi2 = min((hi_ | 31) + T.dhi + 1 + T.sib_off + 31, Lbytes - 1);
					for (int w = (lo2 >> 5) + lane; w <= (hi2 >> 5) && w < nw; w += 32)
						sv_K2[st * nwb + w] = sieve_word_main(pb, st, T2.dupi, T2.flt, w, T2.dlo, T2.dhi,
							[](int, uint32_t f) -> uint32_t { return f; });
					__syncwarp();
				}
				const int k2off = T.sib_off + 1;
				for (int w = (lo_ >> 5) + lane; w <= (hi_ >> 5) && w < nw; w += 32)
					sv_K[st * nwb + w] = sieve_word_main(pb, st, T.dupi, T.flt, w, T.dlo, T.dhi,
						[&](int d, uint32_t f) -> uint32_t {
							// span offset d: the group ends at start + d, its sibling begins k2off later
							return has_k2 ? f & bits32(K2, min((w << 5) + d + k2off, Lbytes)) : f;
						});
			}
		}
		__syncwarp();
	};
	// the lane's word of pass `pass` over the tile's helix starts (strand-major)
	auto sieve_pass = [&](int pass, int &strand_, int &w_) -> uint32_t {
		const DevSearch &SP = sm_ds[max(A.par.pf_search, 0)]; // (unused by a literal-only sieve)
		const int pz = A.par.pf_z;
		const int nv = (int)(gB - gA); // starts of this tile
		const int it = pass * 32 + lane;
		uint32_t word = 0;
		strand_ = it / sv_nws;
		w_ = 0;
		if (strand_ < A.strands) {
			const int zlo = (strand_ ? Lbytes - H - nv : H) + pz, zhi = zlo + nv - 1;
			w_ = (zlo >> 5) + it % sv_nws;
			if (w_ <= (zhi >> 5)) {
				// a span offset d counts only if for some helix length hl the first
				// interior helix can start right after the 5' strand and the last one
				// can end right before the 3' strand (pf_deep; otherwise kh[0] = all ones)
				const int nhl = deep ? SP.maxlen - SP.minlen + 1 : 1;
				const uint32_t *E = sv_E + strand_ * nwb, *K = sv_K + strand_ * nwb;
				uint32_t kh[4];
#pragma unroll
				for (int j = 0; j < 4; j++)
					kh[j] = j < nhl ? (deep && SP.kid_t >= 0 ? bits32(K, (w_ << 5) + SP.minlen + j + SP.kid_off) : ~0u) : 0u;
				// ends asked about at span offset d: e0 + d + m, m = nhl-1-j for length minlen + j
				const int e0 = (w_ << 5) - SP.minlen - SP.lk_off - (nhl - 1);
				const bool has_lk = deep && SP.lk_t >= 0;
				word = !A.par.sv_helix ? ~0u : TWO ? (kh[0] | kh[1] | kh[2] | kh[3]) :
					sieve_word_main(pb, strand_, SP.dupi, SP.flt, w_, SP.dlo, SP.dhi,
					[&](int d, uint32_t f) -> uint32_t {
						if (!has_lk)
							return f & (kh[0] | kh[1] | kh[2] | kh[3]);
						const int q = max(e0 + d, 0);
						const uint32_t x0 = E[q >> 5], x1 = E[(q >> 5) + 1], x2 = E[(q >> 5) + 2];
						const uint32_t lo_ = __funnelshift_r(x0, x1, q & 31), hi_ = __funnelshift_r(x1, x2, q & 31);
						uint32_t la = 0;
#pragma unroll
						for (int j = 0; j < 4; j++)
							if (j < nhl)
								la |= kh[j] & __funnelshift_r(lo_, hi_, nhl - 1 - j);
						return f & la;
					});
				if (CHAIN)
					word &= bits32(ch_F0 + strand_ * nwb, max((w_ << 5) - pz, 0));
				if (LIT) {
					// literal prefilter as a sieve term: the best literal must begin
					// lit_lmin..lit_lmax nucleotides after the start (= helix start - pz)
					const uint32_t *set = sm_lit + strand_ * nwb;
					uint32_t any = 0;
					for (int l = A.par.lit_lmin; l <= A.par.lit_lmax; l++)
						any |= bits32(set, min(max((w_ << 5) - pz + l, 0), Lbytes));
					word &= any;
				}
				// keep the bits of this tile's own starts
				const int b0 = w_ << 5;
				if (zlo > b0)
					word &= ~0u << (zlo - b0);
				if (zhi < b0 + 31)
					word &= ~0u >> (b0 + 31 - zhi);
			}
		}
		return word;
	};
	// take the lowest survivor out of a sieve word: its start item
	auto sieve_pop = [&](uint32_t &word, int strand_, int w_) -> int {
		const int b = (w_ << 5) + __ffs(word) - 1 - A.par.pf_z; // start, in strand buffer coordinates
		word &= word - 1;
		return strand_ ? TILE + (Lbytes - 1 - b - H) : b - H;
	};

	if (MODE == 1) {
		// prefilter only: append the survivors to the global worklist
		auto wl_append = [&](bool pass, int q, uint64_t v0, int have_v0) {
			const unsigned pm = __ballot_sync(0xffffffffu, pass);
			if (pm == 0)
				return;
			unsigned long long base = 0;
			const int leader = __ffs(pm) - 1;
			if (lane == leader) {
				base = atomicAdd(A.wl_count, (unsigned long long)__popc(pm));
				// statistics for the host: entries over all segments (it sizes the
				// segments from the survivor rate) and whether any did not fit
				atomicAdd(A.wl_count + 3, (unsigned long long)__popc(pm));
				if (base + __popc(pm) > A.wl_cap)
					atomicOr(A.wl_count + 4, 1ull);
			}
			base = __shfl_sync(0xffffffffu, base, leader);
			if (pass) {
				const unsigned long long slot = base + __popc(pm & ((1u << lane) - 1));
				if (slot < A.wl_cap) {
					int comp, idx, slen, szero;
					uint32_t rec;
					locate(q, comp, idx, rec, slen, szero);
					const int64_t g = lo + idx;
					uint4 *e = reinterpret_cast<uint4 *>(A.wl + slot * GM_WL_WORDS);
					e[0] = make_uint4((uint32_t)g, (uint32_t)(g >> 32) | ((uint32_t)comp << 31) |
						((uint32_t)have_v0 << 30), rec, (uint32_t)slen);
					e[1] = make_uint4((uint32_t)szero, (uint32_t)v0, (uint32_t)(v0 >> 32), 0u);
				}
			}
		};
		for (;;) {
			const int64_t t = next_tile();
			if (t < 0)
				break;
			load_tile(0, t);
			if (SIEVE) {
				if (CHAIN)
					chain_build();
				if (deep)
					sieve_aux();
				int qhead = 0, qtail = 0; // two-stage sieve: stage-1 survivors, compacted (warp-uniform)
				for (int pass_ = 0; pass_ < sv_npass; pass_++) {
					int st_, w_;
					uint32_t word = sieve_pass(pass_, st_, w_);
					if (ST2) {
						// stage 1 left its survivors in `word`: compact them into the warp's
						// queue and run stage 2 on 32 of them at a time, one per lane
						while (__ballot_sync(0xffffffffu, word != 0)) {
							const bool has = word != 0;
							const int q = has ? sieve_pop(word, st_, w_) : 0;
							const unsigned pm = __ballot_sync(0xffffffffu, has);
							if (has)
								myq[(qtail + __popc(pm & ((1u << lane) - 1))) & (GM_QCAP - 1)] = (uint16_t)q;
							qtail += __popc(pm);
							__syncwarp();
							while (qtail - qhead >= 32) {
								const int q2 = myq[(qhead + lane) & (GM_QCAP - 1)];
								qhead += 32;
								__syncwarp();
								wl_append(accept2(q2), q2, 0, 0);
							}
						}
						continue;
					}
					while (__ballot_sync(0xffffffffu, word != 0)) {
						int q = 0;
						bool pass = false;
						if (word != 0) {
							q = sieve_pop(word, st_, w_);
							pass = accept(q);
						}
						wl_append(pass, q, 0, 0);
					}
				}
				if (qtail != qhead) {
					const bool mine = lane < qtail - qhead;
					const int q2 = mine ? myq[(qhead + lane) & (GM_QCAP - 1)] : 0;
					wl_append(mine && accept2(q2), q2, 0, 0);
				}
			} else {
				for (;;) {
					const int chunk = work_next;
					work_next += 32;
					if (chunk >= n_work)
						break;
					const int q = chunk + lane;
					uint64_t v0;
					int have_v0;
					const bool pass = prefilter(q, v0, have_v0);
					wl_append(pass, q, v0, have_v0);
				}
			}
			__syncwarp();
		}
	} else {
		int s = 0, ph = PH_IDLE;
		int qhead = 0, qtail = 0; // warp-uniform
		// level-0 sieve (DevParams::sieve): each lane holds one word of survivors
		constexpr bool sieve = SIEVE;
		int sv_pass = 0, sv_strand = 0, sv_w = 0;
		uint32_t sw = 0;

		// ---- the machine ------------------------------------------------
		for (;;) {
			const unsigned idle = __ballot_sync(0xffffffffu, ph == PH_IDLE);
			if (__popc(idle) >= refill_min || (idle && no_more_tiles)) {
				const int want = __popc(idle);
				// top the queue up: prefilter chunks of 32 start items; when the
				// tile runs dry move on to the next one in the other buffer
				while (qtail - qhead < want) {
					const unsigned sv_left = sieve ? __ballot_sync(0xffffffffu, sw != 0) : 0u;
					if (!have_tile || (sieve ? (sv_left == 0 && sv_pass >= sv_npass) : work_next >= n_work)) {
						if (no_more_tiles)
							break;
						if (qtail != qhead)
							break; // queued starts still point into the current tile
						const int nb = (cur + 1) % NBUF;
						if (__ballot_sync(0xffffffffu, ph != PH_IDLE && mybuf == nb))
							break; // a lane still enumerates in that buffer
						const int64_t t = next_tile();
						if (t < 0) {
							no_more_tiles = true;
							have_tile = false;
							break;
						}
						load_tile(nb, t);
						sv_pass = 0;
						sw = 0;
						if (CHAIN)
							chain_build();
						if (deep)
							sieve_aux();
						continue;
					}
					if (sieve) {
						if (sv_left == 0) {
							sw = sieve_pass(sv_pass, sv_strand, sv_w);
							sv_pass++;
							continue;
						}
						// every lane with a survivor left hands one over
						int q = 0;
						bool pass = false;
						if (sw != 0) {
							q = sieve_pop(sw, sv_strand, sv_w);
							pass = accept(q);
						}
						const unsigned pm = __ballot_sync(0xffffffffu, pass);
						if (pass)
							myq[(qtail + __popc(pm & ((1u << lane) - 1))) & (GM_QCAP - 1)] = (uint16_t)q;
						qtail += __popc(pm);
						__syncwarp();
						continue;
					}
					const int chunk = work_next;
					work_next += 32;
					const int q = chunk + lane;
					uint64_t v0;
					int have_v0;
					const bool pass = prefilter(q, v0, have_v0);
					const unsigned pm = __ballot_sync(0xffffffffu, pass);
					if (pass)
						myq[(qtail + __popc(pm & ((1u << lane) - 1))) & (GM_QCAP - 1)] = (uint16_t)q;
					qtail += __popc(pm);
					__syncwarp();
				}
				// hand queued starts to idle lanes
				const int avail = qtail - qhead;
				const int rank = __popc(idle & ((1u << lane) - 1));
				if (ph == PH_IDLE && rank < avail) {
					const int q = myq[(qhead + rank) & (GM_QCAP - 1)];
					int comp, idx, slen, szero;
					uint32_t rec;
					locate(q, comp, idx, rec, slen, szero);
					L.rec = rec;
					L.slen = slen;
					L.szero = szero;
					L.comp = comp;
					L.seq = 0;
					strand = comp;
					sqbase = comp ? Lbytes - 1 - idx : idx;
					L.sq = (comp ? sm_rc : sm_fwd) + sqbase;
					mybuf = cur;
					mypb = pb;
					// RM_find_motif, src/find_motif.c:184-205
					L_ZD(L, 0) = pk16(0, min(W, slen - szero) - 1);
					s = 0;
					ph = PH_ENTER;
					my_entries++;
				}
				qhead += min(avail, want);
				__syncwarp();
				if (no_more_tiles && qtail == qhead && __all_sync(0xffffffffu, ph == PH_IDLE))
					break;
			}

#define GM_MASK(S, z, clo, n)                                                                          \
	(FULL ? wc_mask_body(mypb, L.sq, strand, sqbase, (S).dupi, (S).flt, (z), (clo), (n))               \
	      : wc_mask(mypb, L.sq, strand, sqbase, (S).dupi, (S).flt, (z), (clo), (n)))
#define GM_TAIL(S, T, s5, s3, hl) tail_feasible(mypb, L.sq, strand, sqbase, (S), (T), (s5), (s3), (hl))
// 64 bits of the base bitset of base x (table 0) from window-relative position p on
#define GM_BBITS(x, p) bits64(mypb.base + ((size_t)(strand * mypb.n_dups) * 4 + (x)) * mypb.nwb, sqbase + (p))
#define GM_FULL FULL
#include "gm_machine_body.inc"
#undef GM_FULL
#undef GM_TAIL
#undef GM_BBITS
#undef GM_MASK
		}
	}

	// (start, strand) pairs searched, for the stats
	for (int o = 16; o > 0; o >>= 1)
		my_starts += __shfl_down_sync(0xffffffffu, my_starts, o);
	if (lane == 0 && my_starts)
		atomicAdd(A.start_count, my_starts);
	my_entries = __reduce_add_sync(0xffffffffu, my_entries);
	if (lane == 0 && my_entries)
		atomicAdd(A.start_count + 3, (unsigned long long)my_entries);
}

// The two entry points of the tile kernel.  The fused kernel (filter + machine) takes
// what registers it needs; the filter kernel of the worklist path is bounded so that two
// 256-thread blocks share an SM (its shared memory allows no more).
template <int MODE, bool FULL, int PF>
__global__ void gm_search_kernel(const ScanArgs A)
{
	gm_search_body<MODE, FULL, PF>(A);
}
template <int PF>
__global__ void __launch_bounds__(256, 2) gm_filter_kernel(const ScanArgs A)
{
	gm_search_body<1, false, PF>(A);
}

// The enumeration kernel of the worklist path: idle lanes take the filter's
// survivors from the global worklist in batches; every lane builds its own
// window (one byte per nucleotide of the searched strand, reverse complement
// included) and its own pair bitsets straight from the packed database, and the
// lanes run the same machine, with the same masks, as the tile kernel.
template <bool FULL>
__global__ void gm_dfs_kernel(const ScanArgs A)
{
	extern __shared__ __align__(16) uint8_t smem_raw[];
	const int tid = threadIdx.x, nt = blockDim.x;
	const int lane = tid & 31;
	const int NS = A.par.n_searches, ND = A.par.n_descr;
	const int W = A.par.w_winsize;
	const int Lc = A.par.halo - W;            // context nucleotides kept on each side
	const int wstride = A.par.win_stride;     // bytes per lane window (odd number of words)
	const int Wtot = W + 2 * Lc;

	StagedPlan sp;
	uint8_t *p = stage_plan(smem_raw, A, tid, nt, sp);
	DevSearch *sm_ds = sp.ds;
	uint8_t *sm_win = p;                                       p += (size_t)nt * wstride;
	uint32_t *sm_bits = reinterpret_cast<uint32_t *>(p);       p += (size_t)nt * A.par.win_bits * 4;
	uint32_t *sm_state = reinterpret_cast<uint32_t *>(p);

	Lane L;
	L.P = sp.pv;
	L.st = sm_state + tid;
	L.nt = nt;
	L.ds = sm_ds;
	L.ps = sp.ps;
	L.NS = NS;
	L.ND = ND;
	L.el_base = NS + A.par.frame_words;
	uint8_t *mywin = sm_win + (size_t)tid * wstride;
	L.sq = mywin + Lc;
	// the lane's own pair bitsets over its window (same layout as a tile's, one
	// strand; the lane stride is an odd number of words, so the lanes of a warp
	// hit different banks): bit Lc + rel <-> window-relative position rel
	PairBits mypb;
	mypb.base = sm_bits + (size_t)tid * A.par.win_bits;
	mypb.nwb = ((Wtot + 31) >> 5) + 3;
	mypb.n_dups = A.par.n_dups;
	L.szero = L.slen = L.comp = 0;
	L.rec = 0;
	L.seq = 0;
	for (int d = 0; FULL && d < ND; d++) {
		unmark(L, d);
		set_cnt(L, d, GM_UNDEF, GM_UNDEF);
	}
	__syncthreads();

	const unsigned long long wl_n = min(*A.wl_count, A.wl_cap);
	int s = 0, ph = PH_IDLE;
	bool exhausted = false;
	const int nwl = mypb.nwb;
	// entries a warp takes at a time: all its idle lanes when the worklist is long;
	// when it is short (a strong sieve) spread it over every warp of the grid --
	// the lanes of a warp diverge, so a warp with 32 survivors runs 32 enumerations
	// one after the other while the rest of the GPU idles
	const unsigned long long n_warps = (unsigned long long)gridDim.x * (blockDim.x >> 5);
	const int per = (int)max(1ull, min(32ull, (wl_n + n_warps - 1) / n_warps));
	const int rmin = min(A.par.dfs_refill, per);
	const unsigned long long gwarp = (unsigned long long)blockIdx.x * (blockDim.x >> 5) + (tid >> 5);
	int taken = 0;

	for (;;) {
		const unsigned idle = __ballot_sync(0xffffffffu, ph == PH_IDLE);
		if (!exhausted && (__popc(idle) >= rmin || idle == 0xffffffffu)) {
			// hand the idle lanes one worklist entry each.  Every lane builds its own
			// window (one byte per nucleotide of the searched strand, reverse
			// complement included, mk_rcmp src/rnamot.c:193-216) and its own pair
			// bitsets straight from the packed database: a batch costs what one
			// window costs
			const int n = min(__popc(idle), per);
			const int rank = __popc(idle & ((1u << lane) - 1));
			unsigned long long mine;
			if (per < 32) {
				// short worklist: entry j of this warp is j * n_warps + its number, so that
				// neighbouring starts (appended together, and expensive together) go to
				// different warps
				mine = (unsigned long long)(taken + rank) * n_warps + gwarp;
				taken += n;
				if ((unsigned long long)taken * n_warps + gwarp >= wl_n)
					exhausted = true;
			} else {
				unsigned long long base = 0;
				if (lane == 0)
					base = atomicAdd(A.wl_head, (unsigned long long)n);
				base = __shfl_sync(0xffffffffu, base, 0);
				if (base + n >= wl_n)
					exhausted = true;
				mine = base + rank;
			}
			if (ph == PH_IDLE && rank < n && mine < wl_n) {
				const uint4 *e = reinterpret_cast<const uint4 *>(A.wl + mine * GM_WL_WORDS);
				const uint4 e0 = e[0], e1v = e[1];
				L.rec = e0.z;
				L.slen = (int)e0.w;
				L.szero = (int)e1v.x;
				L.comp = (int)(e0.y >> 31);
				L.seq = 0;
				const int64_t roff = A.rec_off[L.rec];
				const int slen = L.slen, comp = L.comp, c00 = L.szero - Lc;
				uint32_t *mb = const_cast<uint32_t *>(mypb.base);
				// Lite plans: eight window positions at a time (expand8, gm_tilebits.h) -- two aligned
				// words of the packed database and a funnel shift bring their codes in line; on the
				// complementary strand the same forward word gives the complement bytes in reversed
				// order and the bit-reversed bitset bytes; positions outside the record read as code 0
				// (0x40 on both strands).  Checked against the per-nucleotide loop below on the host
				// (tests/csrc/winbuild_check.cpp).  The full machine keeps the per-nucleotide loop: with
				// this build in it the kernel (96 instead of 124 registers) never finished on
				// pseudoknot plans, for a reason not found (profiles/win_probe.sh).
				const uint32_t *pw = reinterpret_cast<const uint32_t *>(A.packed);
				const int64_t wmax = (A.total_nt >> 3) + 1; // last word that may be read (the buffer has slack)
				uint32_t *win32 = reinterpret_cast<uint32_t *>(mywin);
				for (int w = 0; w < nwl; w++) {
					uint32_t b0 = 0, b1 = 0, b2 = 0, b3 = 0;
					const int i0 = w << 5;
					if (!FULL) {
#pragma unroll
					for (int g = 0; g < 4; g++) {
						const int t0 = i0 + 8 * g;
						if (t0 < Wtot) {
							const int c0 = c00 + t0;
							const int jlo = max(0, -c0), jhi = min(min(8, slen - c0), Wtot - t0);
							uint32_t x = 0;
							if (jhi > jlo) {
								const int64_t f0 = comp ? roff + (slen - 1 - (c0 + 7)) : roff + c0;
								const int64_t wi = f0 >> 3;
								const int sh = (int)(f0 & 7) * 4;
								x = __funnelshift_r(pw[max((int64_t)0, min(wi, wmax))], pw[max((int64_t)0, min(wi + 1, wmax))], sh);
								if (jlo > 0 || jhi < 8) {
									const int nlo = comp ? 8 - jhi : jlo, nhi = comp ? 8 - jlo : jhi;
									x &= (nhi >= 8 ? ~0u : ((1u << (4 * nhi)) - 1u)) & (~0u << (4 * nlo));
								}
							}
							uint32_t f0_, f1_, r0_, r1_, bits;
							expand8(x, f0_, f1_, r0_, r1_, bits);
							if (comp) {
								f0_ = r0_;
								f1_ = r1_;
								bits = __brev(bits);
								if (jlo > 0 || jhi < 8) {
									for (int jj = 0; jj < 8; jj++)
										if (jj < jlo || jj >= jhi) {
											uint32_t &wd = jj < 4 ? f0_ : f1_;
											wd = (wd & ~(0xffu << (8 * (jj & 3)))) | (0x40u << (8 * (jj & 3)));
										}
								}
							}
							win32[t0 >> 2] = f0_;
							if (t0 + 4 < Wtot)
								win32[(t0 >> 2) + 1] = f1_;
							b0 |= (bits & 0xffu) << (8 * g);
							b1 |= ((bits >> 8) & 0xffu) << (8 * g);
							b2 |= ((bits >> 16) & 0xffu) << (8 * g);
							b3 |= (bits >> 24) << (8 * g);
						}
					}
					} else {
					for (int t = 0; t < 32 && i0 + t < Wtot; t++) {
						const int c = c00 + i0 + t; // strand coordinate
						uint8_t v = (uint8_t)(4 << 4);
						if (c >= 0 && c < slen) {
							const int64_t gf = roff + (comp ? slen - 1 - c : c);
							const unsigned byte = A.packed[gf >> 1];
							v = expand_code((byte >> ((gf & 1) * 4)) & 15);
							if (comp)
								v = complement_byte(v);
						}
						mywin[i0 + t] = v;
						const int bc = bcode_of(v);
						b0 |= (uint32_t)(bc == 0) << t;
						b1 |= (uint32_t)(bc == 1) << t;
						b2 |= (uint32_t)(bc == 2) << t;
						b3 |= (uint32_t)(bc == 3) << t;
					}
					}
					mb[0 * nwl + w] = b0;
					mb[1 * nwl + w] = b1;
					mb[2 * nwl + w] = b2;
					mb[3 * nwl + w] = b3;
					// the other tables: unions of the base bitsets
					for (int dd = 1; dd < mypb.n_dups; dd++) {
						const unsigned dup = A.par.dups[dd];
#pragma unroll
						for (int x = 0; x < 4; x++) {
							const unsigned row = dup >> (x * 5);
							mb[(dd * 4 + x) * nwl + w] = ((row & 1u) ? b0 : 0u) | ((row & 2u) ? b1 : 0u) |
								((row & 4u) ? b2 : 0u) | ((row & 8u) ? b3 : 0u);
						}
					}
				}
				const int dl = min(W, L.slen - L.szero) - 1;
				L_ZD(L, 0) = pk16(0, dl);
				s = 0;
				ph = PH_ENTER;
			}
			__syncwarp();
		}
		if (exhausted && __all_sync(0xffffffffu, ph == PH_IDLE))
			break;
#define GM_MASK(S, z, clo, n) wc_mask(mypb, L.sq, 0, Lc, (S).dupi, (S).flt, (z), (clo), (n))
#define GM_TAIL(S, T, s5, s3, hl) tail_feasible(mypb, L.sq, 0, Lc, (S), (T), (s5), (s3), (hl))
#define GM_BBITS(x, p) bits64(mypb.base + (size_t)(x) * mypb.nwb, Lc + (p))
#define GM_FULL FULL
#include "gm_machine_body.inc"
#undef GM_FULL
#undef GM_TAIL
#undef GM_BBITS
#undef GM_MASK
	}
}

} // namespace gm
