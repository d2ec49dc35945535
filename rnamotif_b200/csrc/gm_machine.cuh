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
	PHX_PK_S5 = CL_PK * 8, PHX_PK_S3,
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

struct Smem {
	uint64_t bar;
	int work;          // next chunk of start items of the current tile
	int r_lo;          // first record intersecting the tile
	int one_rec;       // the tile lies inside a single record
	int pad;
	int64_t tile;      // current tile index (broadcast)
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
	uint64_t a0 = ones, a1 = ones, a2 = ones;
	for (int k = 0; k < req; k++) {
		const int x = bcode_of(sq[z + k]);
		uint64_t m = 0;
		if (x < 4)
			m = bits64(sets + x * pb.nwb, sqbase + lo - k);
		if (k == 0) {
			if (first_must)
				a0 = a1 = a2 = m;
			// else: an outermost mispair is always tolerated (ends without 5'
			// pairing, src/find_motif.c:1014-1017): no constraint from k = 0
		} else {
			a2 = (a2 & m) | a1;
			a1 = (a1 & m) | a0;
			a0 &= m;
		}
	}
	return (budget == 0 ? a0 : budget == 1 ? a1 : a2) & ones;
}

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

	// carve shared memory (mirrors smem_need() on the host)
	Smem *sm = reinterpret_cast<Smem *>(smem_raw);
	uint8_t *p = smem_raw + 64;
	uint8_t *sm_stage = p;               p += stage_bytes;
	uint8_t *sm_fwd = p;                 p += Lbytes;
	uint8_t *sm_rc = p;                  p += Lbytes;
	uint32_t *sm_pb = reinterpret_cast<uint32_t *>(p);         p += (((size_t)2 * n_dups * 4 * nwb * 4) + 15) & ~(size_t)15;
	DevSearch *sm_ds = reinterpret_cast<DevSearch *>(p);       p += ((NS * sizeof(DevSearch) + 15) & ~15);
	gm_pairset_t *sm_ps = reinterpret_cast<gm_pairset_t *>(p); p += ((c_plan.n_pairsets * sizeof(gm_pairset_t) + 15) & ~15);
	uint32_t *sm_elmm = reinterpret_cast<uint32_t *>(p);       p += ((ND * 4 + 15) & ~15);
	int64_t *sm_rec = reinterpret_cast<int64_t *>(p);          p += (GM_REC_CACHE + 1) * 8;
	uint16_t *sm_q = reinterpret_cast<uint16_t *>(p);          p += (size_t)(nt >> 5) * GM_QCAP * 2;
	uint32_t *sm_state = reinterpret_cast<uint32_t *>(p);

	// stage the hot plan tables
	for (int i = tid; i < NS * (int)(sizeof(DevSearch) / 4); i += nt)
		reinterpret_cast<uint32_t *>(sm_ds)[i] = reinterpret_cast<const uint32_t *>(c_ds)[i];
	for (int i = tid; i < c_plan.n_pairsets * (int)(sizeof(gm_pairset_t) / 4); i += nt)
		reinterpret_cast<uint32_t *>(sm_ps)[i] = reinterpret_cast<const uint32_t *>(c_plan.pairsets)[i];
	for (int i = tid; i < ND; i += nt)
		sm_elmm[i] = pk16(c_plan.elems[i].minlen, c_plan.elems[i].maxlen);
	if (tid == 0)
		mbar_init(&sm->bar, 1);

	Lane L;
	L.st = sm_state + tid;
	L.nt = nt;
	L.ds = sm_ds;
	L.ps = sm_ps;
	L.NS = NS;
	L.ND = ND;
	L.el_base = NS + c_par.frame_words;
	L.sq = sm_fwd;
	L.szero = L.slen = L.comp = 0;
	L.rec = 0;
	L.seq = 0;
	// never-marked elements read as UNDEF; counters start at UNDEF like
	// SE_init leaves them (src/compile.c:570-571)
	for (int d = 0; d < ND; d++) {
		unmark(L, d);
		set_cnt(L, d, GM_UNDEF, GM_UNDEF);
	}
	PairBits pb;
	pb.base = sm_pb;
	pb.nwb = nwb;
	pb.n_dups = n_dups;
	uint16_t *myq = sm_q + warp * GM_QCAP;
	__syncthreads();

	uint32_t parity = 0;
	unsigned long long my_starts = 0;
	int sqbase = 0, strand = 0; // tile index of the lane's window start, strand buffer

	for (;;) {
		// ---- next tile ------------------------------------------------
		if (tid == 0)
			sm->tile = (int64_t)atomicAdd(A.tile_counter, 1ull);
		__syncthreads();
		const int64_t t = sm->tile;
		if (t >= A.n_tiles)
			break;
		const int64_t gA = A.g_begin + t * (int64_t)TILE;
		const int64_t gB = min(gA + (int64_t)TILE, A.g_end);
		const int64_t lo = gA - H;
		// packed bytes [bs, bs + nbytes) cover nucleotides [lo_c, hi_c)
		const int64_t lo_c = max(lo, (int64_t)0);
		const int64_t hi_c = min(lo + Lbytes, A.total_nt);
		const int64_t bs = (lo_c >> 1) & ~(int64_t)15;
		const uint32_t nbytes = (uint32_t)((((hi_c + 1) >> 1) - bs + 15) & ~(int64_t)15);
		if (tid == 0) {
			mbar_expect_tx(&sm->bar, nbytes);
			tma_bulk_g2s(sm_stage, A.packed + bs, nbytes, &sm->bar);
			// the last record starting at or before gA
			int a = 0, b = A.n_rec; // rec_off[a] <= gA < rec_off[b]
			while (b - a > 1) {
				int m = (a + b) >> 1;
				if (A.rec_off[m] <= gA)
					a = m;
				else
					b = m;
			}
			sm->r_lo = a;
			sm->one_rec = A.rec_off[a + 1] >= gB;
			sm->work = 0;
		}
		__syncthreads();
		{
			// cache the offsets of up to GM_REC_CACHE records from r_lo on
			const int r_lo = sm->r_lo;
			for (int i = tid; i <= GM_REC_CACHE; i += nt) {
				int r = r_lo + i;
				sm_rec[i] = r <= A.n_rec ? A.rec_off[r] : (int64_t)1 << 62;
			}
		}
		mbar_wait(&sm->bar, parity);
		parity ^= 1;
		// expand packed nibbles to one byte per nucleotide, both strands
		for (int i = tid; i < Lbytes; i += nt) {
			const int64_t g = lo + i;
			uint8_t v = (uint8_t)(4 << 4);
			if (g >= 0 && g < A.total_nt) {
				unsigned byte = sm_stage[(g >> 1) - bs];
				v = expand_code((byte >> ((g & 1) * 4)) & 15);
			}
			sm_fwd[i] = v;
			sm_rc[Lbytes - 1 - i] = complement_byte(v);
		}
		__syncthreads();
		// pair bitsets: one ballot per (strand, table, base) and 32 positions
		for (int w = warp; w < nwb; w += (nt >> 5)) {
			const int i = w * 32 + lane;
			const int vf = i < Lbytes ? bcode_of(sm_fwd[i]) : 4;
			const int vr = i < Lbytes ? bcode_of(sm_rc[i]) : 4;
			for (int dd = 0; dd < n_dups; dd++) {
				const unsigned dup = c_par.dups[dd];
				for (int x = 0; x < 4; x++) {
					const unsigned bf = __ballot_sync(0xffffffffu, (dup >> (x * 5 + vf)) & 1u);
					const unsigned br = __ballot_sync(0xffffffffu, (dup >> (x * 5 + vr)) & 1u);
					if (lane == 0) {
						sm_pb[((size_t)(0 * n_dups + dd) * 4 + x) * nwb + w] = bf;
						sm_pb[((size_t)(1 * n_dups + dd) * 4 + x) * nwb + w] = br;
					}
				}
			}
		}
		__syncthreads();

		const int n_work = A.strands * TILE;
		const bool one_rec = sm->one_rec != 0;
		const int64_t rec0_off = sm_rec[0];
		const int rec0_len = (int)(sm_rec[1] - sm_rec[0]);
		int s = 0, ph = PH_IDLE;
		bool exhausted = false;
		int qhead = 0, qtail = 0; // warp-uniform

		// locate start item q: returns false if it is not a start of this scan
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
				rec = (uint32_t)sm->r_lo;
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
				rec = (uint32_t)(a + sm->r_lo);
			} else {
				int a = sm->r_lo, b = A.n_rec;
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

		// ---- the machine ------------------------------------------------
		for (;;) {
			const unsigned idle = __ballot_sync(0xffffffffu, ph == PH_IDLE);
			if (idle) {
				const int want = __popc(idle);
				// top the queue up: prefilter chunks of 32 start items
				while (qtail - qhead < want && !exhausted) {
					int chunk = 0;
					if (lane == 0)
						chunk = atomicAdd(&sm->work, 32);
					chunk = __shfl_sync(0xffffffffu, chunk, 0);
					if (chunk >= n_work) {
						exhausted = true;
						break;
					}
					const int q = chunk + lane;
					int comp, idx, slen, szero;
					uint32_t rec;
					bool pass = locate(q, comp, idx, rec, slen, szero);
					if (pass) {
						my_starts++;
						const DevSearch &S0 = sm_ds[0];
						const uint8_t *sq = comp ? sm_rc + (Lbytes - 1 - idx) : sm_fwd + idx;
						const int base = comp ? Lbytes - 1 - idx : idx;
						const int dl = min(W, slen - szero) - 1;
						if (S0.dupi >= 0 && (S0.flt & 0xff)) {
							// any span end of search 0 at all?
							int fsd, lsd;
							if (S0.kind == K_PK) {
								fsd = dl;
								lsd = 2 * S0.minlen - 1;
							} else {
								fsd = min(dl, S0.maxglen - 1);
								lsd = S0.minglen - 1;
							}
							bool any = false;
							for (int hi = fsd; hi >= lsd && !any; hi -= 64) {
								const int l0 = max(lsd, hi - 63);
								any = wc_mask(pb, sq, comp, base, S0.dupi, S0.flt, 0, l0, hi - l0 + 1) != 0;
							}
							pass = any;
						}
						if (pass && S0.rx5 >= 0 && S0.mm5 == 0 && !c_plan.regex[S0.rx5].eol) {
							// a seq= without '$' that cannot match the longest
							// placement cannot match a shorter one
							pass = rx_match(c_plan.regex[S0.rx5], sq, min(S0.maxlen, dl + 1)) != 0;
						}
					}
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
					// RM_find_motif, src/find_motif.c:184-205
					L_ZD(L, 0) = pk16(0, min(W, slen - szero) - 1);
					s = 0;
					ph = PH_ENTER;
				}
				qhead += min(avail, want);
				__syncwarp();
				if (exhausted && qtail == qhead && __all_sync(0xffffffffu, ph == PH_IDLE))
					break;
			}

			// majority vote over the classes the lanes are waiting in
			const int cls = ph >> 3;
			const unsigned peers = __match_any_sync(0xffffffffu, cls);
			const int vote = cls == CL_IDLE ? 0 : ((__popc(peers) << 4) | cls);
			const int run_cls = __reduce_max_sync(0xffffffffu, vote) & 15;
			if (cls != run_cls)
				continue;

			int wc_stage = 0; // 0: pick the next span end, 1: keep extending the helix
			switch (ph) {
			case PH_IDLE:
				break;

			case PH_WC_RESUME:
				unmark(L, sm_ds[s].d);
				unmark(L, sm_ds[s].d3);
				wc_stage = 1;
				ph = PH_SPAN;
				goto do_span;

			case PH_SS_RESUME:
				unmark(L, sm_ds[s].d);
				ph = PH_SPAN;
				goto do_span;

			case PH_PH_RESUME:
				unmark(L, sm_ds[s].d);
				unmark(L, sm_ds[s].d3);
				ph = PH_SPAN;
				goto do_span;

			case PH_ENTER: {
				// find_motif, src/find_motif.c:245-287
				const DevSearch &S = sm_ds[s];
				const uint32_t zd = L_ZD(L, s);
				const int z = lo16(zd), dl = hi16(zd);
				int sd, lsd;
				if (S.loop) {
					sd = min(dl, z + S.maxglen - 1);
					lsd = z + S.minglen - 1;
				} else
					sd = lsd = dl;
				if (S.kind == K_SS) {
					// find_ss accepts lengths in [minlen, maxlen] only, src/find_motif.c:349
					sd = min(sd, z + S.maxlen - 1);
					lsd = max(lsd, z + S.minlen - 1);
				}
				L_FR(L, s, 0) = pk16(sd + 1, lsd);
				if (S.kind == K_WC || S.kind == K_QU) {
					// no candidate mask yet
					L_FR(L, s, 5) = 0;
					L_FR(L, s, 6) = 0;
				}
				ph = PH_SPAN;
			}
			// fall through
			case PH_SPAN:
			do_span: {
				const DevSearch &S = sm_ds[s];
				const uint32_t zd = L_ZD(L, s);
				const int z = lo16(zd), dl = hi16(zd);
				const uint32_t w0 = L_FR(L, s, 0);
				int sd = lo16(w0) - 1;
				const int lsd = hi16(w0);

				if (S.kind == K_SS) {
					// find_motif's span loop + find_ss (src/find_motif.c:268-280,332-398)
					// fused: run down the span ends to the next one that is accepted
					for (;; sd--) {
						if (sd < lsd) {
							GM_RETURN();
							break;
						}
						const int len = sd - z + 1;
						set_cnt(L, S.d, 0, 0);
						if (S.rx5 >= 0 && !chk_seq5(L, S, z, len))
							continue;
						mark(L, S.d, z, len);
						if (S.next_s >= 0)
							L_ZD(L, S.next_s) = pk16(sd + 1, dl);
						if (S.last) {
							sink(L, A);
							unmark(L, S.d);
							continue;
						}
						L_FR(L, s, 0) = pk16(sd, lsd);
						L_FR(L, s, 1) = pk16(0, PH_SS_RESUME);
						s++;
						ph = PH_ENTER;
						break;
					}
					break;
				}

				if (S.kind == K_WC) {
					// find_motif's span loop + find_wchlx + match_wchlx
					// (src/find_motif.c:268-280,400-463,975-1112) fused: one trip =
					// advance to the next helix candidate of this level, or return
					const int s5 = z;
					int top = sd, clo = hi16(L_FR(L, s, 4));
					uint64_t v = (uint64_t)L_FR(L, s, 5) | ((uint64_t)L_FR(L, s, 6) << 32);
					int s3 = 0, s3lim = 0, hl = 0, mpr = 0, lbpr = 1;
					if (wc_stage) {
						const uint32_t w3 = L_FR(L, s, 3);
						const int f1 = lo16(L_FR(L, s, 1));
						s3 = hi16(L_FR(L, s, 2));
						s3lim = lo16(w3);
						hl = hi16(w3);
						mpr = f1 & 0xff;
						lbpr = (f1 >> 8) & 1;
					}
					for (;;) {
						bool cand = false;
						if (!wc_stage) {
							// next span end from the candidate mask: `top` is the largest
							// span end no chunk has covered; bit j of v <-> clo + j
							while (v == 0) {
								if (top < lsd)
									break;
								clo = max(lsd, top - 63);
								v = wc_mask(pb, L.sq, strand, sqbase, S.dupi, S.flt, z, clo, top - clo + 1);
								top = clo - 1;
							}
							if (v == 0) {
								GM_RETURN();
								break;
							}
							const int j = 63 - __clzll((long long)v);
							v &= ~(1ull << j);
							s3 = clo + j;
							int t3 = s3 - z + 1;
							t3 = (t3 - S.minilen) / 2;
							t3 = min(t3, S.maxlen);
							s3lim = s3 - t3 + 1;
							hl = 0; mpr = 0; lbpr = 1;
							wc_stage = 1;
							// the empty helix (minlen = 0) comes first, src/find_motif.c:986-1006
							cand = S.minlen == 0 && s3 - s5 + 1 <= S.maxilen;
						}
						if (!cand) {
							if (!wx_next(L, S, s5, s3, s3lim, hl, mpr, lbpr)) {
								wc_stage = 0;
								continue;
							}
							// find_wchlx, src/find_motif.c:441-447
							if (s3 - s5 - 2 * hl + 1 > S.maxilen)
								continue;
						}
						// descend into the interior with this helix
						L_FR(L, s, 0) = pk16(top + 1, lsd);
						L_FR(L, s, 1) = pk16(FR1_LO(mpr, lbpr, 0), PH_WC_RESUME);
						L_FR(L, s, 2) = pk16(s5, s3);
						L_FR(L, s, 3) = pk16(s3lim, hl);
						L_FR(L, s, 4) = pk16(0, clo);
						L_FR(L, s, 5) = (uint32_t)v;
						L_FR(L, s, 6) = (uint32_t)(v >> 32);
						set_cnt(L, S.d, mpr, 0);
						set_cnt(L, S.d3, mpr, 0);
						mark(L, S.d, s5, hl);
						mark(L, S.d3, s3 - hl + 1, hl);
						if (S.next_s >= 0)
							L_ZD(L, S.next_s) = pk16(s3 + 1, dl);
						L_ZD(L, s + 1) = pk16(s5 + hl, s3 - hl);
						s++;
						ph = PH_ENTER;
						break;
					}
					break;
				}
				if (S.kind == K_WC || S.kind == K_QU) {
					// span ends come from the candidate mask.  `top` (= sd here) is
					// the largest span end no chunk has covered yet; a chunk is
					// [clo, clo + 63] and bit j of v stands for span end clo + j.
					uint64_t v = (uint64_t)L_FR(L, s, 5) | ((uint64_t)L_FR(L, s, 6) << 32);
					int clo = hi16(L_FR(L, s, 4));
					int top = sd;
					while (v == 0) {
						if (top < lsd)
							break;
						clo = max(lsd, top - 63);
						v = wc_mask(pb, L.sq, strand, sqbase, S.dupi, S.flt, z, clo, top - clo + 1);
						top = clo - 1;
					}
					if (v == 0) {
						L_FR(L, s, 0) = pk16(top + 1, lsd);
						GM_RETURN();
						break;
					}
					const int j = 63 - __clzll((long long)v);
					v &= ~(1ull << j);
					sd = clo + j;
					L_FR(L, s, 5) = (uint32_t)v;
					L_FR(L, s, 6) = (uint32_t)(v >> 32);
					L_FR(L, s, 4) = pk16(0, clo);
					L_FR(L, s, 0) = pk16(top + 1, lsd);
				} else {
					if (sd < lsd) {
						GM_RETURN();
						break;
					}
					L_FR(L, s, 0) = pk16(sd, lsd);
				}
				if (S.next_s >= 0)
					L_ZD(L, S.next_s) = pk16(sd + 1, dl);

				switch (S.kind) {
				case K_SS: {
					// find_ss, src/find_motif.c:332-398
					const int len = sd - z + 1;
					set_cnt(L, S.d, 0, 0);
					if (len < S.minlen || len > S.maxlen)
						break;
					if (S.rx5 >= 0 && !chk_seq5(L, S, z, len))
						break;
					mark(L, S.d, z, len);
					if (S.last) {
						sink(L, A);
						unmark(L, S.d);
					} else {
						L_FR(L, s, 1) = pk16(0, PH_SS_RESUME);
						s++;
						ph = PH_ENTER;
					}
					break;
				}
				case K_WC:
				case K_QU: {
					// find_wchlx :400-433 / find_4plex :851-892
					set_cnt(L, S.d, 0, 0);
					set_cnt(L, S.d3, 0, 0);
					int i_minl = S.minilen;
					if (S.kind == K_QU) {
						const gm_elem_t &e = c_plan.elems[S.d];
						set_cnt(L, e.mates[0], 0, 0);
						set_cnt(L, e.mates[1], 0, 0);
						i_minl = S.minilen + c_plan.elems[e.mates[0]].minilen +
							c_plan.elems[e.mates[1]].minilen + 2 * S.minlen;
					}
					int t3 = sd - z + 1;
					t3 = (t3 - i_minl) / 2;
					t3 = min(t3, S.maxlen);
					const int s3lim = sd - t3 + 1;
					L_FR(L, s, 2) = pk16(z, sd);
					L_FR(L, s, 3) = pk16(s3lim, 0);
					ph = S.minlen == 0 ? PH_WX_BEGIN : PH_WX_FIRST;
					break;
				}
				case K_PK: {
					// find_pknot + find_pknot5, src/find_motif.c:465-528
					const gm_elem_t &e = c_plan.elems[S.d];
					const int *sc = &c_plan.scopes[e.scopes];
					if (e.scope == 0) {
						for (int k = 1; k < e.n_scopes; k++) {
							const int d1 = sc[k];
							if (c_plan.elems[d1].type == GM_H5) {
								unmark(L, d1);
								L_ZD(L, c_plan.elems[d1].searchno) = pk16(z, sd);
							}
						}
					}
					const int d0 = sc[0], dn = sc[e.n_scopes - 1];
					const int slen = sd - z + 1;
					const int p_minl = pk_minlen(L, sm_elmm, d0, S.d - 1);
					const int p_maxl = pk_maxlen(L, sm_elmm, d0, S.d - 1);
					const int r_minl = pk_minlen(L, sm_elmm, S.d, dn);
					const int r_maxl = pk_maxlen(L, sm_elmm, S.d, dn);
					if (p_maxl + r_maxl < slen)
						break;
					const int f_s5 = z + p_minl;
					const int l_s5 = z + min(p_maxl, slen - r_minl);
					L_FR(L, s, 4) = pk16(f_s5 - 1, l_s5);
					ph = PH_PK_S5;
					break;
				}
				case K_PH: {
					// find_phlx, src/find_motif.c:703-761
					set_cnt(L, S.d, 0, 0);
					set_cnt(L, S.d3, 0, 0);
					const int slen = sd - z + 1;
					int s5hi = min((slen - S.minilen) / 2, S.maxlen);
					s5hi = z + s5hi - 1;
					int ilen = slen - 2 * S.minlen;
					ilen = min(ilen, S.maxilen);
					int s5lo = slen - ilen;
					if (s5lo & 1)
						s5lo++;
					s5lo = min(s5lo / 2, S.maxlen);
					s5lo = z + s5lo - 1;
					int hlen, n_mpr;
					if (!match_phlx(L, S, S.d3, z, sd, s5hi, s5lo, &hlen, &n_mpr))
						break;
					if (sd - z - 2 * hlen + 1 > S.maxilen)
						break;
					set_mpr(L, S.d, n_mpr);
					set_mpr(L, S.d3, n_mpr);
					mark(L, S.d, z, hlen);
					mark(L, S.d3, sd - hlen + 1, hlen);
					L_ZD(L, s + 1) = pk16(z + hlen, sd - hlen);
					L_FR(L, s, 1) = pk16(0, PH_PH_RESUME);
					s++;
					ph = PH_ENTER;
					break;
				}
				case K_TR: {
					// find_triplex, src/find_motif.c:763-819
					const gm_elem_t &e = c_plan.elems[S.d];
					const int dd1 = e.mates[0], dd2 = e.mates[1];
					const gm_elem_t &e1 = c_plan.elems[dd1];
					set_cnt(L, S.d, 0, 0);
					set_cnt(L, dd1, 0, 0);
					set_cnt(L, dd2, 0, 0);
					const int slen = sd - z + 1;
					int s5hi = min((slen - S.minilen - e1.minilen) / 2, S.maxlen);
					s5hi = z + s5hi - 1;
					int i_len = slen - 2 * S.minlen;
					i_len = min(i_len, S.maxilen + S.minlen + e1.maxilen);
					int s5lo = slen - i_len;
					if (s5lo & 1)
						s5lo++;
					s5lo = min(s5lo / 2, S.maxlen);
					s5lo = z + s5lo - 1;
					int hlen, n_mpr;
					if (!match_phlx(L, S, dd2, z, sd, s5hi, s5lo, &hlen, &n_mpr))
						break;
					if (sd - z - 2 * hlen + 1 > S.maxilen + e1.maxilen + hlen)
						break;
					mark(L, S.d, z, hlen);
					mark(L, dd2, sd - hlen + 1, hlen);
					L_FR(L, s, 2) = pk16(z, sd);
					L_FR(L, s, 4) = pk16(sd - e1.minilen - hlen + 1, hlen);
					ph = PH_TR_S;
					break;
				}
				}
				break;
			}

			case PH_WX_BEGIN: {
				// the empty-helix candidate of match_wchlx, src/find_motif.c:986-1006
				// (gm_plan_check refuses seq= on a minlen=0 helix, so it is unconditional)
				const DevSearch &S = sm_ds[s];
				const uint32_t w2 = L_FR(L, s, 2);
				const int s5 = lo16(w2), s3 = hi16(w2);
				L_FR(L, s, 1) = pk16(FR1_LO(0, 1, 0), PH_WX_RESUME);
				// after this candidate the first pair is tested: hl stays 0
				if (S.kind == K_WC) {
					if (s3 - s5 + 1 > S.maxilen) {
						ph = PH_WX_FIRST;
						break;
					}
					set_mpr(L, S.d, 0);
					set_mpr(L, S.d3, 0);
					mark(L, S.d, s5, 0);
					mark(L, S.d3, s3 + 1, 0);
					L_ZD(L, s + 1) = pk16(s5, s3);
					s++;
					ph = PH_ENTER;
				} else if (S.kind == K_QU) {
					mark(L, S.d, s5, 0);
					mark(L, S.d3, s3 + 1, 0);
					L_FR(L, s, 7) = pk16(s5 + S.minilen - 1, 0);
					ph = PH_QU_S1;
				} else
					ph = PH_WX_FIRST; // K_PK with minlen 0 is refused by gm_plan_check
				break;
			}

			case PH_WX_RESUME: {
				const DevSearch &S = sm_ds[s];
				unmark(L, S.d);
				unmark(L, S.d3);
				if (hi16(L_FR(L, s, 3)) == 0) {
					// came back from the empty-helix candidate
					ph = PH_WX_FIRST;
					break;
				}
				ph = PH_WX_EXT;
			}
			// fall through
			case PH_WX_FIRST:
			case PH_WX_EXT: {
				// match_wchlx, src/find_motif.c:1008-1109, one candidate at a time
				const DevSearch &S = sm_ds[s];
				const uint32_t w2 = L_FR(L, s, 2), w3 = L_FR(L, s, 3);
				const int s5 = lo16(w2), s3 = hi16(w2), s3lim = lo16(w3);
				int hl, mpr, lbpr, chk;
				if (ph == PH_WX_FIRST) {
					if (paired(S.duplex, L.sq[s5], L.sq[s3])) {
						hl = 1; mpr = 0; lbpr = 1;
					} else if (!(S.ends & GM_5PAIRED)) {
						hl = 1; mpr = 1; lbpr = 0;
					} else {
						ph = S.kind == K_PK ? PH_PK_S3 : PH_SPAN;
						break;
					}
					chk = 1;
				} else {
					const int f1 = lo16(L_FR(L, s, 1));
					hl = hi16(w3);
					mpr = f1 & 0xff;
					lbpr = (f1 >> 8) & 1;
					chk = (f1 >> 9) & 1;
				}
				int found = 0;
				for (;;) {
					if (chk) {
						chk = 0;
						if (hl >= S.minlen &&
						    !(!lbpr && (S.ends & GM_3PAIRED)) &&
						    !(S.pfrac && mpr > c_plan.lentab[S.lentab + hl]) &&
						    !(S.rx5 >= 0 && !rx_match(c_plan.regex[S.rx5], L.sq + s5, hl)) &&
						    !(S.rx3 >= 0 && !rx_match(c_plan.regex[S.rx3], L.sq + s3 - hl + 1, hl))) {
							if (S.kind == K_WC) {
								// find_wchlx, src/find_motif.c:441-447
								if (s3 - s5 - 2 * hl + 1 <= S.maxilen)
									found = 1;
							} else if (S.kind == K_PK) {
								const int i_minl = hi16(L_FR(L, s, 5));
								// find_pknot3, src/find_motif.c:609-627
								if ((s3 - s5 + 1) - 2 * hl < i_minl) {
									found = -1; // "break": no more for this s3
									break;
								}
								found = 1;
								const gm_elem_t &e = c_plan.elems[S.d];
								const int *sc = &c_plan.scopes[e.scopes];
								if (S.d == sc[1]) {
									const int d3_h1 = c_plan.elems[sc[0]].mates[0];
									const int iL_last = m_off(L, d3_h1) - 1;
									const int iR_last = m_off(L, d3_h1) + m_len(L, d3_h1);
									int iL_minl = 0, iL_maxl = 0, iR_minl = 0, iR_maxl = 0;
									if (S.d + 1 <= d3_h1 - 1) {
										iL_minl = pk_minlen(L, sm_elmm, S.d + 1, d3_h1 - 1);
										iL_maxl = pk_maxlen(L, sm_elmm, S.d + 1, d3_h1 - 1);
									}
									if (d3_h1 + 1 <= S.d3 - 1) {
										iR_minl = pk_minlen(L, sm_elmm, d3_h1 + 1, S.d3 - 1);
										iR_maxl = pk_maxlen(L, sm_elmm, d3_h1 + 1, S.d3 - 1);
									}
									const int iL = iL_last - (s5 + hl - 1);
									const int iR = (s3 - hl + 1) - iR_last;
									if (iL < iL_minl || iL > iL_maxl || iR < iR_minl || iR > iR_maxl)
										found = 0;
								}
							} else
								found = 1; // K_QU: every helix goes to find_4plex_inner
							if (found)
								break;
						}
					}
					if (s3 - hl + 1 < s3lim || hl >= S.maxlen) {
						found = -1;
						break;
					}
					if (paired(S.duplex, L.sq[s5 + hl], L.sq[s3 - hl]))
						lbpr = 1;
					else {
						if (++mpr > S.mplim) {
							found = -1;
							break;
						}
						lbpr = 0;
					}
					hl++;
					chk = 1;
				}
				if (found < 0) {
					ph = S.kind == K_PK ? PH_PK_S3 : PH_SPAN;
					break;
				}
				// a candidate: remember where the extension stands
				L_FR(L, s, 3) = pk16(s3lim, hl);
				L_FR(L, s, 1) = pk16(FR1_LO(mpr, lbpr, 0), PH_WX_RESUME);
				mark(L, S.d, s5, hl);
				mark(L, S.d3, s3 - hl + 1, hl);
				if (S.kind == K_WC) {
					set_mpr(L, S.d, mpr);
					set_mpr(L, S.d3, mpr);
					L_ZD(L, s + 1) = pk16(s5 + hl, s3 - hl);
					s++;
					ph = PH_ENTER;
				} else if (S.kind == K_PK) {
					set_mpr(L, S.d, mpr);
					set_mpr(L, S.d3, mpr);
					upd_pksearches(L, S.d, s5, s3, hl);
					s++;
					ph = PH_ENTER;
				} else {
					// find_4plex_inner, src/find_motif.c:937-938
					L_FR(L, s, 7) = pk16(s5 + hl + S.minilen - 1, 0);
					ph = PH_QU_S1;
				}
				break;
			}

			case PH_PK_S5: {
				// find_pknot5 loop + find_pknot3 prologue, src/find_motif.c:523-568
				const DevSearch &S = sm_ds[s];
				const uint32_t w4 = L_FR(L, s, 4);
				const int s5 = lo16(w4) + 1, l_s5 = hi16(w4);
				if (s5 > l_s5) {
					ph = PH_SPAN;
					break;
				}
				L_FR(L, s, 4) = pk16(s5, l_s5);
				const gm_elem_t &e = c_plan.elems[S.d];
				const int dn = c_plan.scopes[e.scopes + e.n_scopes - 1];
				const int sd = lo16(L_FR(L, s, 0));
				const int slen = sd - s5 + 1;
				const int i_minl = pk_minlen(L, sm_elmm, S.d + 1, S.d3 - 1);
				const int g_minl = 2 * S.minlen + i_minl;
				const int s_minl = pk_minlen(L, sm_elmm, S.d3 + 1, dn);
				const int s_maxl = pk_maxlen(L, sm_elmm, S.d3 + 1, dn);
				if (g_minl + s_minl > slen)
					break; // next s5
				const int f_s3 = sd - s_minl;
				const int l_s3 = sd - min(slen - g_minl, s_maxl);
				L_FR(L, s, 2) = pk16(s5, f_s3 + 1);
				L_FR(L, s, 5) = pk16(l_s3, i_minl);
				L_FR(L, s, 7) = 0;
				L_FR(L, s, 8) = 0;
				L_FR(L, s, 6) = pk16(f_s3, 0); // (largest 3' end not yet covered, chunk low end)
				ph = PH_PK_S3;
				break;
			}

			case PH_PK_S3: {
				// find_pknot3 loop over the 3' end, src/find_motif.c:600-606,
				// through the same candidate mask as the proper helices
				const DevSearch &S = sm_ds[s];
				const uint32_t w2 = L_FR(L, s, 2), w5 = L_FR(L, s, 5), w6 = L_FR(L, s, 6);
				const int s5 = lo16(w2);
				const int l_s3 = lo16(w5), i_minl = hi16(w5);
				int top = lo16(w6), clo = hi16(w6);
				uint64_t v = (uint64_t)L_FR(L, s, 7) | ((uint64_t)L_FR(L, s, 8) << 32);
				while (v == 0) {
					if (top < l_s3)
						break;
					clo = max(l_s3, top - 63);
					v = wc_mask(pb, L.sq, strand, sqbase, S.dupi, S.flt, s5, clo, top - clo + 1);
					top = clo - 1;
				}
				if (v == 0) {
					ph = PH_PK_S5;
					break;
				}
				const int j = 63 - __clzll((long long)v);
				v &= ~(1ull << j);
				const int s3 = clo + j;
				L_FR(L, s, 7) = (uint32_t)v;
				L_FR(L, s, 8) = (uint32_t)(v >> 32);
				L_FR(L, s, 6) = pk16(top, clo);
				L_FR(L, s, 2) = pk16(s5, s3);
				int t3 = s3 - s5 + 1;
				t3 = (t3 - i_minl) / 2;
				t3 = min(t3, S.maxlen);
				L_FR(L, s, 3) = pk16(s3 - t3 + 1, 0);
				ph = PH_WX_FIRST;
				break;
			}

			case PH_TR_RESUME:
				unmark(L, c_plan.elems[sm_ds[s].d].mates[0]);
				ph = PH_TR_S;
			// fall through
			case PH_TR_S: {
				// find_triplex loop over the t2 end, src/find_motif.c:821-845
				const DevSearch &S = sm_ds[s];
				const gm_elem_t &e = c_plan.elems[S.d];
				const int dd1 = e.mates[0], dd2 = e.mates[1];
				const gm_elem_t &e1 = c_plan.elems[dd1];
				const uint32_t w4 = L_FR(L, s, 4), w2 = L_FR(L, s, 2);
				const int sp = lo16(w4) - 1, hlen = hi16(w4);
				const int z = lo16(w2), sd = hi16(w2);
				if (sp < z + 2 * hlen + S.minilen - 1) {
					unmark(L, S.d);
					unmark(L, dd2);
					ph = PH_SPAN;
					break;
				}
				L_FR(L, s, 4) = pk16(sp, hlen);
				int n_mpr;
				if (!match_triplex(L, S, dd1, z, sp, sd, hlen, &n_mpr))
					break;
				if (sp - 2 * hlen - z + 1 > S.maxilen)
					break;
				if (sd - hlen - sp > e1.maxilen)
					break;
				set_mpr(L, S.d, n_mpr);
				set_mpr(L, dd1, n_mpr);
				set_mpr(L, dd2, n_mpr);
				mark(L, dd1, sp - hlen + 1, hlen);
				L_ZD(L, s + 1) = pk16(z + hlen, sp - hlen);
				L_ZD(L, c_plan.elems[e1.inner].searchno) = pk16(sp + 1, sd - hlen);
				L_FR(L, s, 1) = pk16(0, PH_TR_RESUME);
				s++;
				ph = PH_ENTER;
				break;
			}

			case PH_QU_S1: {
				// find_4plex_inner outer loop, src/find_motif.c:937-939
				const DevSearch &S = sm_ds[s];
				const gm_elem_t &e = c_plan.elems[S.d];
				const int i2_minl = c_plan.elems[e.mates[0]].minilen;
				const int i3_minl = c_plan.elems[e.mates[1]].minilen;
				const int s3 = hi16(L_FR(L, s, 2)), hl = hi16(L_FR(L, s, 3));
				const int s1 = lo16(L_FR(L, s, 7)) + 1;
				if (s1 > s3 - 3 * hl - i3_minl - i2_minl) {
					// this q1/q4 helix is done: back to the extension
					unmark(L, S.d);
					unmark(L, S.d3);
					ph = hl == 0 ? PH_WX_FIRST : PH_WX_EXT;
					break;
				}
				L_FR(L, s, 7) = pk16(s1, s3 - hl - i3_minl + 1);
				ph = PH_QU_S2;
				break;
			}

			case PH_QU_RESUME: {
				const gm_elem_t &e = c_plan.elems[sm_ds[s].d];
				unmark(L, e.mates[0]);
				unmark(L, e.mates[1]);
				ph = PH_QU_S2;
			}
			// fall through
			case PH_QU_S2: {
				// find_4plex_inner inner loop, src/find_motif.c:940-969
				const DevSearch &S = sm_ds[s];
				const gm_elem_t &e = c_plan.elems[S.d];
				const int dd1 = e.mates[0], dd2 = e.mates[1];
				const gm_elem_t &e1 = c_plan.elems[dd1], &e2 = c_plan.elems[dd2];
				const uint32_t w2 = L_FR(L, s, 2), w7 = L_FR(L, s, 7);
				const int z = lo16(w2), s3 = hi16(w2), hl = hi16(L_FR(L, s, 3));
				const int s1 = lo16(w7), s2 = hi16(w7) - 1;
				if (s2 < s1 + 2 * hl + e1.minilen) {
					ph = PH_QU_S1;
					break;
				}
				L_FR(L, s, 7) = pk16(s1, s2);
				int n_mpr;
				if (!match_4plex(L, dd1, dd2, z, s1, s2, s3, hl, &n_mpr))
					break;
				if (s1 - z - hl + 1 > S.maxilen)
					break;
				if (s2 - s1 - 2 * hl + 1 > e1.maxilen)
					break;
				if (s3 - s2 - hl + 1 > e2.maxilen)
					break;
				set_mpr(L, S.d, n_mpr);
				set_mpr(L, dd1, n_mpr);
				set_mpr(L, dd2, n_mpr);
				set_mpr(L, S.d3, n_mpr);
				mark(L, dd1, s1, hl);
				mark(L, dd2, s2 - hl + 1, hl);
				L_ZD(L, s + 1) = pk16(z + hl, s1 - 1);
				L_ZD(L, c_plan.elems[e1.inner].searchno) = pk16(s1 + hl, s2 - hl);
				L_ZD(L, c_plan.elems[e2.inner].searchno) = pk16(s2 + 1, s3 - hl);
				L_FR(L, s, 1) = pk16(lo16(L_FR(L, s, 1)), PH_QU_RESUME);
				s++;
				ph = PH_ENTER;
				break;
			}

			}
		}
		__syncthreads(); // everyone is done with this tile's shared memory
	}

	// (start, strand) pairs searched, for the stats
	for (int o = 16; o > 0; o >>= 1)
		my_starts += __shfl_down_sync(0xffffffffu, my_starts, o);
	if (lane == 0 && my_starts)
		atomicAdd(A.start_count, my_starts);
}

} // namespace gm
