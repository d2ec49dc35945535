// gm_search.cuh -- the per-start-position descriptor search on the device.
//
// One lane runs one (start offset, strand) at a time through an explicit-stack
// enumeration that visits candidates in exactly the order of the reference's
// recursion (src/find_motif.c:245-973).  The recursion there always descends
// from search s to search s+1 (find_ss -> s_forward, find_wchlx/phlx/triplex/
// 4plex -> the first interior search, find_pknot3 -> searchno+1;
// gm_plan_check verifies this for every plan), so the stack is an array of
// per-search frames indexed by s and "return" means s-1.
//
// All coordinates are relative to the lane's start offset (szero == 0), which
// keeps every field in 16 bits; the reference's absolute coordinates differ
// by the constant szero everywhere except the places noted at the sink.
//
// Lanes of a warp are at different depths most of the time; the machine is
// therefore written as ONE loop whose body is a switch over a small set of
// phases, so that lanes in the same phase -- whatever their depth -- execute
// together.  Lane state lives in shared memory as 32-bit words laid out
// [word][thread] (bank = thread: conflict-free for any mix of depths).
#pragma once

#include <stddef.h>
#include <stdint.h>
#include "gpumotif_plan.h"

// The bitsets and the lane state live in shared memory, but where their pointers have gone
// through a struct or a call the compiler no longer knows that and emits generic loads with
// 64-bit address arithmetic (430 LD.E in the two-stage filter kernel, none after).  Telling it
// turns them into LDS with 32-bit addresses: fewer instructions (-19 % in the enumeration
// kernel's SASS), fewer registers (the filter kernel no longer spills).  Used ONLY on the
// bitset readers (bits64 / bits32 / sieve_word_r) and the lane-state accessors: wrapping the
// staged plan tables or the sequence bytes as well -- also shared memory -- gave wrong
// candidates with nvcc 12.9 (bisected on the GPU: profiles/bisect_sh.sh), so those stay generic.
template <typename T>
__device__ __forceinline__ T *gm_sh(T *p)
{
	__builtin_assume(__isShared((const void *)p));
	return p;
}

namespace gm {

// kinds of search heads (what find_1_motif dispatches on, src/find_motif.c:289-330)
enum { K_SS = 0, K_WC = 1, K_PK = 2, K_PH = 3, K_TR = 4, K_QU = 5 };

// hot per-search parameters, derived on the host from gm_plan_t
struct DevSearch {
	int kind;
	int d;          // head element
	int d3;         // far strand: h3 / p3 / t3 / q4 (mates[last]); -1 for ss
	int loop;       // find_motif: iterate the span end (src/find_motif.c:255-264)
	int next_s;     // search of s_next, or -1
	int minlen, maxlen;
	int minglen, maxglen;
	int minilen, maxilen;
	int ends, pfrac, mplim;
	unsigned duplex;       // ps_mat[0] as 25 bits
	int lentab;            // per-length pairfrac table (offset into plan.lentab) or -1
	int rx5, rx3;          // regex of head / far strand, or -1
	int mm5;               // mismatch limit of the head (ss: used by the fused ss trip)
	int hmm;               // helix: d or d3 has seq= with mismatch= (their counters need the
	                       // whole extension of match_wchlx, see wx_finish_mm)
	int dupi_t;            // index of the TRANSPOSED duplex table in DevParams::dups (reverse masks), or -1
	int last;              // 1 for the final search (hit sink follows)
	int fr;                // word offset of this search's frame in the lane state
	int dupi;              // index of `duplex` in DevParams::dups (pair bitsets), or -1
	int flt;               // span-end prefilter: req | budget << 8 | first_must << 16
	// look-ahead pruning, in DESCRIPTOR order (elements are contiguous on the sequence):
	// the next helix head reachable through fixed-length single strands from the end of
	// this helix's 5' strand (kid) and from the end of its 3' strand (sib), with the
	// nucleotides in between; -1 if there is none or it is searched earlier
	int kid_t, kid_off;
	int sib_t, sib_off;
	// tail look-ahead: the helix whose 3' strand ends lk_off nucleotides (fixed-length
	// single strands) before this helix's 3' strand begins: its 3' end is known as
	// soon as this helix is chosen
	int lk_t, lk_off;
	// Span offsets s3 - s5 the helix can take at all: [minglen - 1, maxglen - 1] for a
	// proper helix or quadruplex (what find_motif's span loop visits); for a pseudoknot
	// helix the static bounds from the lengths of the elements between its strands.
	int dlo, dhi;
	int nest;               // bit 0: kid_t lies inside this helix (its span ends before s3 - hl)
	// Probes: single strands of fixed length with seq= whose place is known as soon as
	// this helix is chosen (only fixed-length single strands between them and one of
	// the helix's boundaries) and which are searched later: if one cannot match there,
	// the subtree cannot reach the hit sink.
	// packed: anchor (2 bits: 0 after the 5' strand, 1 after the 3' strand, 2 before the
	// 3' strand) | off << 2 (10 bits) | len << 12 (8) | regex << 20 (5) | mismatch << 25 (4)
	int n_probe;
	unsigned probe[4];
	// pseudoknot helix: the elements of its pseudoknot that can be matched when the
	// search reaches it (the strands of the helices searched before it), as a slice of
	// DevParams::pk_m -- find_minlen / find_maxlen are prefix sums corrected for those
	int pkm_off, pkm_n;
};
static_assert(sizeof(DevSearch) == 41 * 4, "DevSearch is staged with an odd word stride");

#define GM_MAX_DUPS 8
#define GM_MAX_CHAIN 40
#define GM_MAX_PKM 64

struct DevParams {
	int n_searches, n_descr;
	int n_pairsets, n_regex, n_scopes, n_lentab, n_sites; // used prefixes of the plan's tables
	int w_winsize;          // min(rm_dmaxlen, windowsize), src/find_motif.c:179
	int dminlen;
	int strict_helices;
	int halo;               // nucleotides staged on each side of a tile
	int tile;               // starts per tile
	int words_per_lane;
	int frame_words;        // sum of all frames
	int win_stride;         // split path: bytes of one lane window
	int win_stage;          // split path: packed bytes staged per worklist entry (multiple of 16)
	int win_bits;           // split path: words of one lane's pair bitsets (odd)
	int n_dups;             // distinct duplex tables with pair bitsets
	unsigned dups[GM_MAX_DUPS];
	int refill_min;         // idle lanes a warp waits for before it hands out new starts
	int dfs_refill;         // the same in the enumeration kernel, where every refill builds lane windows
	                        // (measured, 512 Mnt: pk_j1+2 12.1 -> 13.9 at 16 instead of 4, trna.general 20.2 -> 22.3 at 24 instead of 8)
	int pf_search;          // search whose candidate mask is the level-0 prefilter, or -1
	int pf_z;               // its 5' start relative to the window (fixed-length ss before it)
	int sieve;              // level-0 sieve (sieve_word): word-parallel test of pf_search's span ends
	int pf_deep;            // second stage behind the sieve: kid / tail look-ahead per span end
	int sv_helix;           // the sieve has a helix term (pf_search); otherwise only the literal term
	int sv_two;             // two-stage sieve: stage 1 = the look-ahead bitsets (first interior helix and
	                        // its sibling), chain and literal terms as words; stage 2 = the first helix's
	                        // span-end test per surviving START (wc_mask) -- cheaper than the word-parallel
	                        // main pass when stage 1 leaves few starts
	int sv_id;              // index in dups[] of the identity table (its bitsets are the base bitsets)
	// Composition chain (a term of the level-0 sieve, chain_build in gm_machine.cuh): the
	// descriptor's elements from last to first as steps (min, max, allowed bases,
	// exception budget per length); a start survives only if every element can be laid
	// out contiguously behind it with its strand made of bases that can pair at all.
	int pk_m[GM_MAX_PKM];   // see DevSearch::pkm_off
	int chain;              // number of steps, 0 = off
	unsigned chain_w0[GM_MAX_CHAIN]; // min (12 bits) | max << 12 (12) | allowed bases << 24 (4) | both ends << 28 | constrained << 29
	unsigned chain_w1[GM_MAX_CHAIN]; // budget of length min + i in bits 2i, 2i+1 (3 = no limit)
	int lit_present;        // literal prefilter (gm_plan_t::literal): regex index, window, length
	int lit_rx, lit_lmin, lit_lmax, lit_mm, lit_len;
	int lite;               // plan of single strands and proper helices only: the lane
	                        // state has no per-element counter words; the sink reads the
	                        // mispair / mismatch counts from the frames through elsrc[]
	int elsrc[GM_MAX_DESCR]; // lite: search whose frame word 1 holds element d's count
};

// frame words: 0 (sd, lsd)  1 (flags, resume phase)  2 (s5, s3)  3 (s3lim, hl)
//              4 kind-specific  5,6 span-end candidate mask (PK: 5 = (l_s3, i_minl), 7,8 = mask)
#define GM_FW_SS 2
#define GM_FW_HX 7
#define GM_FW_QU 9
#define GM_FW_PK 18 // 9..15: find_minlen/find_maxlen results, constant while the level is active; 16,17: candidate mask of the level

__device__ __forceinline__ uint32_t pk16(int a, int b)
{
	return (uint32_t)(a & 0xffff) | ((uint32_t)b << 16);
}
__device__ __forceinline__ int lo16(uint32_t w) { return (int)(short)(w & 0xffff); }
__device__ __forceinline__ int hi16(uint32_t w) { return (int)w >> 16; }

// tile byte: low nibble = IUPAC code, high nibble = reference base code 0..4
__device__ __forceinline__ int bcode_of(int v) { return v >> 4; }
__device__ __forceinline__ int icode_of(int v) { return v & 15; }

// RM_paired, src/find_motif.c:1291-1302
__device__ __forceinline__ int paired(unsigned duplex, int v5, int v3)
{
	return (duplex >> (bcode_of(v5) * 5 + bcode_of(v3))) & 1u;
}

// The plan as the kernels see it: every table a kernel reads with a per-lane index is
// staged into SHARED memory at kernel start (stage_plan, gm_kernel.cuh) from the
// context's own device copy of the plan.  (It used to live in __constant__ memory:
// one symbol per device, shared by every context on it, and a divergent index into
// constant memory is replayed once per distinct address.)
struct DevRegex {           // the part of gm_regex_t the device reads (its first 176 bytes)
	int32_t n_items, bol, eol, npos, closure_iters, mm_len;
	uint64_t skip, star, dot;
	uint64_t B[16];
};
static_assert(sizeof(DevRegex) == offsetof(gm_regex_t, items), "DevRegex mirrors the head of gm_regex_t");

struct PlanView {
	DevParams par;
	const gm_elem_t *elems;
	const gm_pairset_t *pairsets;
	const DevRegex *regex;
	const int32_t *scopes;
	const uint8_t *lentab;
	const gm_site_t *sites;
	gm_ctxel_t lctx, rctx;
	int n_sites, n_pairsets, n_regex, pad_;
};

struct Lane {
	// views into shared memory
	const PlanView *P;       // the staged plan
	uint32_t *st;            // this thread's column of the state array
	int nt;                  // threads per block (row stride)
	const uint8_t *sq;       // sq[rel] = tile byte at window-relative position
	const DevSearch *ds;     // staged search table
	const gm_pairset_t *ps;  // staged pairsets
	int NS, ND;
	int el_base;             // first word of the per-element state
	// the start this lane is working on
	int szero;               // absolute offset in the searched strand
	int slen;                // record length
	int comp;
	uint32_t rec;
	uint32_t seq;            // candidates emitted so far for this start
};

// ---- lane state accessors -------------------------------------------------
#define L_ZD(L, s)     gm_sh((L).st)[(s) * (L).nt]
#define L_FR(L, s, k)  gm_sh((L).st)[((L).NS + gm_sh((L).ds)[s].fr + (k)) * (L).nt]
#define L_EL(L, d)     gm_sh((L).st)[((L).el_base + (d)) * (L).nt]
#define L_EM(L, d)     gm_sh((L).st)[((L).el_base + (L).ND + (d)) * (L).nt]

__device__ __forceinline__ void mark(Lane &L, int d, int off, int len)
{
	L_EL(L, d) = pk16(off, len);
}
__device__ __forceinline__ void unmark(Lane &L, int d)
{
	L_EL(L, d) = pk16(GM_UNDEF, GM_UNDEF);
}
// (s_matchoff, s_matchlen) of element d as one packed word.  Plans with
// pseudoknots etc. keep them in the lane state (find_minlen/find_maxlen read
// them during the search); "lite" plans do not store them at all -- they are
// only needed at the hit sink, where they follow from the frames.
__device__ uint32_t el_word_lite(const Lane &L, int d);
__device__ __forceinline__ uint32_t el_word(const Lane &L, int d, bool lite)
{
	return lite ? el_word_lite(L, d) : L_EL(L, d);
}
__device__ __forceinline__ void set_cnt(Lane &L, int d, int mpr, int mm)
{
	L_EM(L, d) = pk16(mpr, mm);
}
__device__ __forceinline__ void set_mpr(Lane &L, int d, int mpr)
{
	L_EM(L, d) = pk16(mpr, hi16(L_EM(L, d)));
}

// ---- regex: bit-parallel NFA (step/advance, src/regexp.c:389-664) -----------
// Boolean result only: for the operator subset the plan admits (classes, '.',
// '*', \{m,n\}, '^', '$') "some backtracking path succeeds" is regular-language
// membership, which the position automaton decides exactly.
__device__ __noinline__ int rx_match(const DevRegex &rx_, const uint8_t *s_, int n)
{
	const DevRegex &rx = rx_;
	const uint8_t *s = s_;
	const uint64_t skip = rx.skip, star = rx.star;
	const uint64_t accept = (uint64_t)1 << rx.npos;
	const int iters = rx.closure_iters;
	uint64_t a0 = 1;
	for (int i = 0; i < iters; i++)
		a0 |= (a0 & skip) << 1;
	uint64_t a = a0;
	bool any = (a & accept) != 0;
	const bool bol = rx.bol != 0, eol = rx.eol != 0;
	if (!eol && any)
		return 1;
	for (int j = 0; j < n; j++) {
		uint64_t m = a & rx.B[icode_of(s[j])];
		a = (m << 1) | (m & star);
		for (int i = 0; i < iters; i++)
			a |= (a & skip) << 1;
		if (!bol)
			a |= a0;
		if (a & accept) {
			if (!eol)
				return 1;
		}
		if (a == 0)
			return 0;
	}
	return (a & accept) != 0 ? 1 : 0;
}

// mm_step/mm_advance, src/mm_regexp.c:353-469 (fixed-length patterns).
// Returns 1 and the mismatch count of the first (leftmost) placement that
// stays within l_mm; on failure *n_mm is what the last placement tried left
// behind (mm_advance counts into the caller's s_n_mismatches as it goes).
__device__ __noinline__ int rx_match_mm(const DevRegex &rx_, const uint8_t *s_, int n, int l_mm, int *n_mm)
{
	const DevRegex &rx = rx_;
	const uint8_t *s = s_;
	const int m = rx.mm_len;
	const int last = rx.bol ? 0 : n;
	int cnt = 0;
	for (int p1 = 0; p1 <= last; p1++) {
		cnt = 0;
		bool ok = true;
		for (int k = 0; k < m; k++) {
			if (p1 + k >= n) { ok = false; break; }
			uint64_t bit = (uint64_t)1 << k;
			if (!(rx.dot & bit) && !(rx.B[icode_of(s[p1 + k])] & bit)) {
				if (++cnt > l_mm) { ok = false; break; }
			}
		}
		if (ok && rx.eol && p1 + m != n)
			ok = false;
		if (ok) {
			*n_mm = cnt;
			return 1;
		}
	}
	*n_mm = cnt;
	return 0;
}

// ---- multi-strand pair rules ---------------------------------------------------
// RM_triple / RM_quad, src/find_motif.c:1304-1331
__device__ __forceinline__ int triple(const gm_pairset_t &p, int v1, int v2, int v3)
{
	int k = (bcode_of(v1) * 5 + bcode_of(v2)) * 5 + bcode_of(v3);
	return (p.multi[k >> 5] >> (k & 31)) & 1u;
}
__device__ __forceinline__ int quad(const gm_pairset_t &p, int v1, int v2, int v3, int v4)
{
	int k = ((bcode_of(v1) * 5 + bcode_of(v2)) * 5 + bcode_of(v3)) * 5 + bcode_of(v4);
	return (p.multi[k >> 5] >> (k & 31)) & 1u;
}

} // namespace gm
