// gm_machine.cuh -- gm_search_kernel: tile staging, span-end prefilter, and
// the search machine.
//
// Per tile (TILE consecutive nucleotides of the concatenated database, both
// strands):
//   1. one TMA bulk copy brings the packed tile + halo into shared memory;
//      it is expanded to one byte per nucleotide for the forward strand and
//      for the reverse complement (mk_rcmp, src/rnamot.c:193-216);
//   2. "pair bitsets" are built with __ballot_sync: for every distinct duplex
//      table D used by a helix head and every base x, bit p of P[D][x] says
//      whether x pairs with the nucleotide at tile position p.  With them the
//      set of span ends whose outermost `req` pairs can form is a handful of
//      funnel shifts and ANDs (wc_mask) instead of a loop over span ends;
//   3. warps pull chunks of 32 start positions; a start whose level-0 mask is
//      empty (or whose anchored seq= cannot match) is dropped at once, the
//      others go to a small per-warp queue;
//   4. lanes take starts from the queue and run the explicit-stack machine.
//      The masks are hints that only ever remove span ends match_wchlx would
//      reject at its first tests (src/find_motif.c:1010-1079); every survivor
//      goes through the full test, so the enumeration is unchanged.
#pragma once

#include "gm_kernel.cuh"

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
__device__ __forceinline__ uint64_t bits64(const uint32_t *set, int q)
{
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
__device__ uint64_t wc_mask(const PairBits &pb, const uint8_t *sq, int strand, int sqbase,
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
// LIT = true adds the literal prefilter (plans with a gm_plan_t::literal); a
// template parameter so that plans without one run exactly the code they had
// before it existed.
template <int MODE, bool FULL, bool LIT>
__global__ void gm_search_kernel(const ScanArgs A)
{
	extern __shared__ __align__(16) uint8_t smem_raw[];
	const int tid = threadIdx.x, nt = blockDim.x;
	const int lane = tid & 31, warp = tid >> 5;
	const int NS = c_par.n_searches, ND = c_par.n_descr;
	const int W = c_par.w_winsize, H = c_par.halo, TILE = c_par.tile;
	const int Lbytes = (TILE + 2 * H + 15) & ~15;          // nucleotides staged per tile
	const int stage_bytes = ((Lbytes >> 1) + 32 + 15) & ~15; // packed staging (+ alignment slack)
	const int nwb = ((Lbytes + 31) >> 5) + 4;
	const int n_dups = c_par.n_dups;
	const int NBUF = MODE == 0 ? 2 : 1;

	// carve shared memory (mirrors smem_need() on the host): plan tables shared
	// by the block, then one private region per warp, then the lane state
	const int nwarps = nt >> 5;
	const size_t pb_bytes = (((size_t)2 * n_dups * 4 * nwb * 4) + 15) & ~(size_t)15;
	const size_t lit_bytes = LIT ? (((size_t)2 * nwb * 4) + 15) & ~(size_t)15 : 0;
	const size_t buf_bytes = 2 * (size_t)Lbytes + pb_bytes + (GM_REC_CACHE + 2) * 8 + lit_bytes;
	const size_t warp_bytes = 16 + (size_t)stage_bytes + NBUF * buf_bytes + GM_QCAP * 2;
	uint8_t *p = smem_raw;
	DevSearch *sm_ds = reinterpret_cast<DevSearch *>(p);       p += ((NS * sizeof(DevSearch) + 15) & ~15);
	gm_pairset_t *sm_ps = reinterpret_cast<gm_pairset_t *>(p); p += ((c_plan.n_pairsets * sizeof(gm_pairset_t) + 15) & ~15);
	uint32_t *sm_elmm = reinterpret_cast<uint32_t *>(p);       p += ((ND * 4 + 15) & ~15);
	uint64_t *sm_litB = reinterpret_cast<uint64_t *>(p);       p += LIT ? 16 * 8 : 0; // literal prefilter class masks
	uint8_t *wp = p + (size_t)warp * warp_bytes;               p += (size_t)nwarps * warp_bytes;
	uint32_t *sm_state = reinterpret_cast<uint32_t *>(p);
	WarpTile *sm = reinterpret_cast<WarpTile *>(wp);           wp += 16;
	uint8_t *sm_stage = wp;                                    wp += stage_bytes;
	uint8_t *bufs = wp;                                        wp += NBUF * buf_bytes;
	uint16_t *myq = reinterpret_cast<uint16_t *>(wp);

	// stage the hot plan tables
	for (int i = tid; i < NS * (int)(sizeof(DevSearch) / 4); i += nt)
		reinterpret_cast<uint32_t *>(sm_ds)[i] = reinterpret_cast<const uint32_t *>(c_ds)[i];
	for (int i = tid; i < c_plan.n_pairsets * (int)(sizeof(gm_pairset_t) / 4); i += nt)
		reinterpret_cast<uint32_t *>(sm_ps)[i] = reinterpret_cast<const uint32_t *>(c_plan.pairsets)[i];
	for (int i = tid; i < ND; i += nt)
		sm_elmm[i] = pk16(c_plan.elems[i].minlen, c_plan.elems[i].maxlen);
	if (LIT && tid < 16)
		sm_litB[tid] = c_plan.regex[c_par.lit_rx].B[tid];
	if (lane == 0)
		mbar_init(&sm->bar, 1);

	Lane L;
	L.st = sm_state + tid;
	L.nt = nt;
	L.ds = sm_ds;
	L.ps = sm_ps;
	L.NS = NS;
	L.ND = ND;
	L.el_base = NS + c_par.frame_words;
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
	const int refill_min = c_par.refill_min;

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
		// expand packed nibbles to one byte per nucleotide, both strands
		for (int i = lane; i < Lbytes; i += 32) {
			const int64_t g = lo + i;
			uint8_t v = (uint8_t)(4 << 4);
			if (g >= 0 && g < A.total_nt) {
				unsigned byte = sm_stage[(g >> 1) - bs];
				v = expand_code((byte >> ((g & 1) * 4)) & 15);
			}
			sm_fwd[i] = v;
			sm_rc[Lbytes - 1 - i] = complement_byte(v);
		}
		__syncwarp();
		// pair bitsets: one ballot per (strand, table, base) and 32 positions
		uint32_t *pbw = const_cast<uint32_t *>(pb.base);
		for (int w = 0; w < nwb; w++) {
			const int i = w * 32 + lane;
			const int vf = i < Lbytes ? bcode_of(sm_fwd[i]) : 4;
			const int vr = i < Lbytes ? bcode_of(sm_rc[i]) : 4;
			for (int dd = 0; dd < n_dups; dd++) {
				const unsigned dup = c_par.dups[dd];
				for (int x = 0; x < 4; x++) {
					const unsigned bf = __ballot_sync(0xffffffffu, (dup >> (x * 5 + vf)) & 1u);
					const unsigned br = __ballot_sync(0xffffffffu, (dup >> (x * 5 + vr)) & 1u);
					if (lane == 0) {
						pbw[((size_t)(0 * n_dups + dd) * 4 + x) * nwb + w] = bf;
						pbw[((size_t)(1 * n_dups + dd) * 4 + x) * nwb + w] = br;
					}
				}
			}
		}
		if (LIT) {
			// literal prefilter (adjust_szero, src/find_motif.c:209-243): bit i of a
			// strand's set = the best literal occurs at tile position i within its
			// mismatch allowance (mm_advance on fixed-length items, src/mm_regexp.c:369-469)
			const int len = c_par.lit_len, l_mm = c_par.lit_mm;
			const uint64_t dot = c_plan.regex[c_par.lit_rx].dot;
			for (int w = 0; w < nwb; w++) {
				const int i = w * 32 + lane;
				bool mf = i + len <= Lbytes, mr = mf;
				int cf = 0, cr = 0;
				for (int k = 0; k < len && (mf || mr); k++) {
					const uint64_t bit = (uint64_t)1 << k;
					if (dot & bit)
						continue;
					if (mf && !(sm_litB[icode_of(sm_fwd[i + k])] & bit) && ++cf > l_mm)
						mf = false;
					if (mr && !(sm_litB[icode_of(sm_rc[i + k])] & bit) && ++cr > l_mm)
						mr = false;
				}
				const unsigned bf = __ballot_sync(0xffffffffu, mf);
				const unsigned br = __ballot_sync(0xffffffffu, mr);
				if (lane == 0) {
					sm_lit[w] = bf;
					sm_lit[nwb + w] = br;
				}
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
		return slen - szero >= c_par.dminlen;
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
			const int l = c_par.lit_lmin;
			int h = min(c_par.lit_lmax, dl + 1 - c_par.lit_len);
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
		if (c_par.pf_search >= 0) {
			// any span end at all for the first helix of the descriptor?  (It is
			// search 0, or follows fixed-length single strands, so its 5' start
			// pz is known.)
			const DevSearch &SP = sm_ds[c_par.pf_search];
			const int pz = c_par.pf_z;
			int fsd, lsd;
			if (SP.kind == K_PK) {
				fsd = dl;
				lsd = pz + 2 * SP.minlen - 1;
			} else {
				fsd = min(dl, pz + SP.maxglen - 1);
				lsd = pz + SP.minglen - 1;
			}
			bool any = false;
			for (int hi = fsd; hi >= lsd && !any; hi -= 64) {
				const int l0 = max(lsd, hi - 63);
				const uint64_t v = wc_mask(pb, sq, comp, base, SP.dupi, SP.flt, pz, l0, hi - l0 + 1);
				any = v != 0;
				if (hi == fsd && l0 == lsd && SP.kind != K_PK && c_par.pf_search == 0) {
					v0 = v;
					have_v0 = 1;
				}
			}
			pass = any;
		}
		if (pass && S0.rx5 >= 0 && S0.mm5 == 0 && !c_plan.regex[S0.rx5].eol) {
			// a seq= without '$' that cannot match the longest
			// placement cannot match a shorter one
			pass = rx_match(c_plan.regex[S0.rx5], sq, min(S0.maxlen, dl + 1)) != 0;
		}
		return pass;
	};

	if (MODE == 1) {
		// prefilter only: append the survivors to the global worklist
		for (;;) {
			const int64_t t = next_tile();
			if (t < 0)
				break;
			load_tile(0, t);
			for (;;) {
				const int chunk = work_next;
				work_next += 32;
				if (chunk >= n_work)
					break;
				const int q = chunk + lane;
				uint64_t v0;
				int have_v0;
				const bool pass = prefilter(q, v0, have_v0);
				const unsigned pm = __ballot_sync(0xffffffffu, pass);
				if (pm == 0)
					continue;
				unsigned long long base = 0;
				const int leader = __ffs(pm) - 1;
				if (lane == leader)
					base = atomicAdd(A.wl_count, (unsigned long long)__popc(pm));
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
			}
			__syncwarp();
		}
	} else {
		int s = 0, ph = PH_IDLE;
		int qhead = 0, qtail = 0; // warp-uniform

		// ---- the machine ------------------------------------------------
		for (;;) {
			const unsigned idle = __ballot_sync(0xffffffffu, ph == PH_IDLE);
			if (__popc(idle) >= refill_min || (idle && no_more_tiles)) {
				const int want = __popc(idle);
				// top the queue up: prefilter chunks of 32 start items; when the
				// tile runs dry move on to the next one in the other buffer
				while (qtail - qhead < want) {
					if (!have_tile || work_next >= n_work) {
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
				}
				qhead += min(avail, want);
				__syncwarp();
				if (no_more_tiles && qtail == qhead && __all_sync(0xffffffffu, ph == PH_IDLE))
					break;
			}

#define GM_MASK(S, z, clo, n) wc_mask(mypb, L.sq, strand, sqbase, (S).dupi, (S).flt, (z), (clo), (n))
#define GM_FULL FULL
#include "gm_machine_body.inc"
#undef GM_FULL
#undef GM_MASK
		}
	}

	// (start, strand) pairs searched, for the stats
	for (int o = 16; o > 0; o >>= 1)
		my_starts += __shfl_down_sync(0xffffffffu, my_starts, o);
	if (lane == 0 && my_starts)
		atomicAdd(A.start_count, my_starts);
}

// First-pair scan used where no pair bitsets exist (gm_dfs_kernel): span ends
// in [lo, lo+n) whose outermost pair can form -- what match_wchlx tests first
// (src/find_motif.c:1010-1021).  A superset filter like wc_mask.
__device__ __forceinline__ uint64_t wc_mask_scan(const Lane &L, const DevSearch &S, int z, int lo, int n)
{
	const uint64_t ones = n >= 64 ? ~0ull : ((1ull << n) - 1);
	if (!((S.flt >> 16) & 1) || S.minlen == 0)
		return ones;
	const unsigned row = S.duplex >> (bcode_of(L.sq[z]) * 5);
	uint64_t v = 0;
	for (int j = 0; j < n; j++)
		v |= (uint64_t)((row >> bcode_of(L.sq[lo + j])) & 1u) << j;
	return v;
}

struct DfsSmem {
	int dummy[16];
};

// The worklist consumer of the split path: lanes take prefilter survivors from
// the global worklist (32 at a time per warp), the warp cooperatively builds
// each new lane's private window (one byte per nucleotide of the searched
// strand, reverse complement included) straight from the packed database, and
// the lanes run the same machine as the tile kernel.
template <bool FULL>
__global__ void gm_dfs_kernel(const ScanArgs A)
{
	extern __shared__ __align__(16) uint8_t smem_raw[];
	const int tid = threadIdx.x, nt = blockDim.x;
	const int lane = tid & 31, warp = tid >> 5;
	const int NS = c_par.n_searches, ND = c_par.n_descr;
	const int W = c_par.w_winsize;
	const int Lc = c_par.halo - W;            // context nucleotides kept on each side
	const int wstride = c_par.win_stride;     // bytes per lane window (odd number of words)
	const int Wtot = W + 2 * Lc;

	uint8_t *p = smem_raw;
	DevSearch *sm_ds = reinterpret_cast<DevSearch *>(p);       p += ((NS * sizeof(DevSearch) + 15) & ~15);
	gm_pairset_t *sm_ps = reinterpret_cast<gm_pairset_t *>(p); p += ((c_plan.n_pairsets * sizeof(gm_pairset_t) + 15) & ~15);
	uint32_t *sm_elmm = reinterpret_cast<uint32_t *>(p);       p += ((ND * 4 + 15) & ~15);
	uint32_t *sm_ent = reinterpret_cast<uint32_t *>(p);        p += (size_t)nt * GM_WL_WORDS * 4;
	uint8_t *sm_wst = p;                                       p += (size_t)nt * c_par.win_stage;
	uint8_t *sm_win = p;                                       p += (size_t)nt * wstride;
	uint32_t *sm_state = reinterpret_cast<uint32_t *>(p);

	for (int i = tid; i < NS * (int)(sizeof(DevSearch) / 4); i += nt)
		reinterpret_cast<uint32_t *>(sm_ds)[i] = reinterpret_cast<const uint32_t *>(c_ds)[i];
	for (int i = tid; i < c_plan.n_pairsets * (int)(sizeof(gm_pairset_t) / 4); i += nt)
		reinterpret_cast<uint32_t *>(sm_ps)[i] = reinterpret_cast<const uint32_t *>(c_plan.pairsets)[i];
	for (int i = tid; i < ND; i += nt)
		sm_elmm[i] = pk16(c_plan.elems[i].minlen, c_plan.elems[i].maxlen);

	Lane L;
	L.st = sm_state + tid;
	L.nt = nt;
	L.ds = sm_ds;
	L.ps = sm_ps;
	L.NS = NS;
	L.ND = ND;
	L.el_base = NS + c_par.frame_words;
	uint8_t *mywin = sm_win + (size_t)tid * wstride;
	L.sq = mywin + Lc;
	L.szero = L.slen = L.comp = 0;
	L.rec = 0;
	L.seq = 0;
	for (int d = 0; FULL && d < ND; d++) {
		unmark(L, d);
		set_cnt(L, d, GM_UNDEF, GM_UNDEF);
	}
	uint32_t *went = sm_ent + (size_t)warp * 32 * GM_WL_WORDS;
	const int wst = c_par.win_stage;          // packed bytes staged per worklist entry
	uint8_t *wstage = sm_wst + (size_t)warp * 32 * wst;
	__syncthreads();

	const unsigned long long wl_n = min(*A.wl_count, A.wl_cap);
	int s = 0, ph = PH_IDLE;
	bool exhausted = false;
	int bavail = 0, bnext = 0; // entries of the current batch not yet handed out (warp-uniform)

	for (;;) {
		const unsigned idle = __ballot_sync(0xffffffffu, ph == PH_IDLE);
		if (idle) {
			if (bavail == 0 && !exhausted) {
				// next batch of 32 worklist entries, staged in shared memory
				unsigned long long base = 0;
				if (lane == 0)
					base = atomicAdd(A.wl_head, 32ull);
				base = __shfl_sync(0xffffffffu, base, 0);
				if (base >= wl_n)
					exhausted = true;
				else {
					bavail = (int)min((unsigned long long)32, wl_n - base);
					bnext = 0;
					if (lane < bavail) {
						const uint4 *e = reinterpret_cast<const uint4 *>(A.wl + (base + lane) * GM_WL_WORDS);
						uint4 *d = reinterpret_cast<uint4 *>(went + lane * GM_WL_WORDS);
						const uint4 e0 = e[0], e1v = e[1];
						d[0] = e0;
						d[1] = e1v;
						// stage the packed bytes under this entry's window now, all 32
						// entries of the batch at once (one memory latency per batch)
						const int ecomp = (int)(e0.y >> 31), eslen = (int)e0.w, eszero = (int)e1v.x;
						const int64_t roff = A.rec_off[e0.z];
						int c0 = max(eszero - Lc, 0), c1 = min(eszero - Lc + Wtot, eslen); // strand coords [c0, c1)
						if (c1 < c0)
							c1 = c0;
						const int64_t gs = roff + (ecomp ? eslen - c1 : c0); // first forward nucleotide
						const int64_t b16 = (gs >> 1) & ~(int64_t)15;
						d[1].w = (uint32_t)((gs >> 1) - b16) | ((uint32_t)(gs & 1) << 8); // byte/nibble of gs in the stage
						const uint4 *src = reinterpret_cast<const uint4 *>(A.packed + b16);
						uint4 *dst = reinterpret_cast<uint4 *>(wstage + (size_t)lane * wst);
						for (int k = 0; k < (wst >> 4); k++)
							dst[k] = src[k];
					}
					__syncwarp();
				}
			}
			const int rank = __popc(idle & ((1u << lane) - 1));
			const bool take = ph == PH_IDLE && rank < bavail;
			const unsigned tm = __ballot_sync(0xffffffffu, take);
			uint32_t e1 = 0, v0lo = 0, v0hi = 0;
			if (take) {
				const uint32_t *e = went + (bnext + rank) * GM_WL_WORDS;
				e1 = e[1];
				L.rec = e[2];
				L.slen = (int)e[3];
				L.szero = (int)e[4];
				L.comp = (int)(e1 >> 31);
				L.seq = 0;
				v0lo = e[5];
				v0hi = e[6];
			}
			// build the windows of the lanes that took a start, one lane at a time,
			// all 32 lanes copying
			const int myslot = bnext + rank; // batch slot of the entry this lane took
			for (unsigned m = tm; m; m &= m - 1) {
				const int j = __ffs(m) - 1;
				const int jcomp = __shfl_sync(0xffffffffu, L.comp, j);
				const int jslen = __shfl_sync(0xffffffffu, L.slen, j);
				const int jszero = __shfl_sync(0xffffffffu, L.szero, j);
				const int jslot = __shfl_sync(0xffffffffu, myslot, j);
				const uint32_t where = went[jslot * GM_WL_WORDS + 7];
				const int nib0 = (int)(where & 0xff) * 2 + (int)((where >> 8) & 1); // nibble index of gs
				const uint8_t *stg = wstage + (size_t)jslot * wst;
				uint8_t *win = sm_win + (size_t)((warp << 5) + j) * wstride;
				const int c0 = max(jszero - Lc, 0), c1 = max(min(jszero - Lc + Wtot, jslen), c0);
				for (int i = lane; i < Wtot; i += 32) {
					const int c = jszero - Lc + i; // strand coordinate
					uint8_t v = (uint8_t)(4 << 4);
					if (c >= c0 && c < c1) {
						// forward nucleotide gs + k lies k nibbles into the stage
						const int k = nib0 + (jcomp ? c1 - 1 - c : c - c0);
						const unsigned byte = stg[k >> 1];
						v = expand_code((byte >> ((k & 1) * 4)) & 15);
						if (jcomp)
							v = complement_byte(v);
					}
					win[i] = v;
				}
			}
			__syncwarp();
			if (take) {
				const DevSearch &S0 = sm_ds[0];
				const int dl = min(W, L.slen - L.szero) - 1;
				L_ZD(L, 0) = pk16(0, dl);
				s = 0;
				ph = PH_ENTER;
				if ((e1 >> 30) & 1) {
					// the prefilter already computed search 0's candidate mask
					const int lsd = S0.minglen - 1;
					L_FR(L, 0, 0) = pk16(lsd, lsd);   // nothing left above the chunk
					L_FR(L, 0, 4) = pk16(0, lsd);
					L_FR(L, 0, 5) = v0lo;
					L_FR(L, 0, 6) = v0hi;
					ph = PH_SPAN;
				}
			}
			const int took = __popc(tm);
			bavail -= took;
			bnext += took;
			if (exhausted && bavail == 0 && __all_sync(0xffffffffu, ph == PH_IDLE))
				break;
		}
#define GM_MASK(S, z, clo, n) wc_mask_scan(L, (S), (z), (clo), (n))
#define GM_FULL FULL
#include "gm_machine_body.inc"
#undef GM_FULL
#undef GM_MASK
	}
}

} // namespace gm
