// gm_machine.cuh -- gm_search_kernel: tile staging + the search machine.
#pragma once

#include "gm_kernel.cuh"

namespace gm {

// resume phases: what a level does when its child has been exhausted
enum {
	PH_SS_RESUME = 16, PH_WX_RESUME, PH_PH_RESUME, PH_TR_RESUME, PH_QU_RESUME
};

// frame word 3, low half: mpr (8 bits) | l_bpr << 8 | chk << 9
#define FR3_LO(mpr, lbpr, chk) (((mpr) & 0xff) | ((lbpr) << 8) | ((chk) << 9))

struct Smem {
	uint64_t bar;
	int work;          // next work item of the current tile
	int r_lo, r_n;     // records intersecting the tile: first index, count (0 = use global table)
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

__global__ void gm_search_kernel(const ScanArgs A)
{
	extern __shared__ __align__(16) uint8_t smem_raw[];
	const int tid = threadIdx.x, nt = blockDim.x;
	const int lane = tid & 31;
	const int NS = c_par.n_searches, ND = c_par.n_descr;
	const int W = c_par.w_winsize, H = c_par.halo, TILE = c_par.tile;
	const int Lbytes = (TILE + 2 * H + 15) & ~15;          // nucleotides staged per tile
	const int stage_bytes = ((Lbytes >> 1) + 32 + 15) & ~15; // packed staging (+ alignment slack)

	// carve shared memory
	Smem *sm = reinterpret_cast<Smem *>(smem_raw);
	uint8_t *p = smem_raw + 64;
	uint8_t *sm_stage = p;               p += stage_bytes;
	uint8_t *sm_fwd = p;                 p += Lbytes;
	uint8_t *sm_rc = p;                  p += Lbytes;
	DevSearch *sm_ds = reinterpret_cast<DevSearch *>(p);       p += ((NS * sizeof(DevSearch) + 15) & ~15);
	gm_pairset_t *sm_ps = reinterpret_cast<gm_pairset_t *>(p); p += ((c_plan.n_pairsets * sizeof(gm_pairset_t) + 15) & ~15);
	uint32_t *sm_elmm = reinterpret_cast<uint32_t *>(p);       p += ((ND * 4 + 15) & ~15);
	int64_t *sm_rec = reinterpret_cast<int64_t *>(p);          p += (GM_REC_CACHE + 1) * 8;
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
	__syncthreads();

	uint32_t parity = 0;
	unsigned long long my_starts = 0;

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
			// records intersecting [gA, gB): binary search for the one holding gA
			int a = 0, b = A.n_rec; // rec_off[a] <= gA < rec_off[b]
			while (b - a > 1) {
				int m = (a + b) >> 1;
				if (A.rec_off[m] <= gA)
					a = m;
				else
					b = m;
			}
			// skip empty records that start exactly at gA
			while (a + 1 < A.n_rec && A.rec_off[a + 1] <= gA)
				a++;
			sm->r_lo = a;
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

		const int n_work = A.strands * TILE;
		int s = 0, ph = PH_IDLE;
		bool exhausted = false;

		// ---- the machine ------------------------------------------------
		for (;;) {
			const unsigned idle = __ballot_sync(0xffffffffu, ph == PH_IDLE);
			if (idle) {
				if (idle == 0xffffffffu && exhausted)
					break;
				if (!exhausted) {
					int base = 0;
					const int leader = __ffs(idle) - 1;
					if (lane == leader)
						base = atomicAdd(&sm->work, __popc(idle));
					base = __shfl_sync(0xffffffffu, base, leader);
					if (base + __popc(idle) >= n_work)
						exhausted = true;
					if (ph == PH_IDLE) {
						const int q = base + __popc(idle & ((1u << lane) - 1));
						if (q < n_work) {
							const int comp = q >= TILE;
							const int64_t g = gA + (comp ? q - TILE : q);
							if (g < gB) {
								// record holding g
								int a = 0, b = GM_REC_CACHE;
								int64_t off, nxt;
								if (g < sm_rec[GM_REC_CACHE]) {
									while (b - a > 1) {
										int m = (a + b) >> 1;
										if (sm_rec[m] <= g)
											a = m;
										else
											b = m;
									}
									off = sm_rec[a];
									nxt = sm_rec[a + 1];
									a += sm->r_lo;
								} else {
									a = sm->r_lo;
									b = A.n_rec;
									while (b - a > 1) {
										int m = (a + b) >> 1;
										if (A.rec_off[m] <= g)
											a = m;
										else
											b = m;
									}
									off = A.rec_off[a];
									nxt = A.rec_off[a + 1];
								}
								const int slen = (int)(nxt - off);
								const int pos = (int)(g - off);
								const int szero = comp ? slen - 1 - pos : pos;
								const int avail = slen - szero;
								if (avail >= c_par.dminlen && c_par.dminlen > 0) {
									L.rec = (uint32_t)a;
									L.slen = slen;
									L.szero = szero;
									L.comp = comp;
									L.seq = 0;
									const int idx = (int)(g - lo);
									L.sq = comp ? sm_rc + (Lbytes - 1 - idx) : sm_fwd + idx;
									// RM_find_motif, src/find_motif.c:184-205
									L_ZD(L, 0) = pk16(0, min(W, avail) - 1);
									s = 0;
									ph = PH_ENTER;
									my_starts++;
								}
							}
						}
					}
				}
			}

			switch (ph) {
			case PH_IDLE:
				break;

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
				L_FR(L, s, 0) = pk16(sd + 1, lsd);
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
				if (sd < lsd) {
					ph = PH_RET;
					break;
				}
				if (S.kind == K_WC && S.minlen > 0 && (S.ends & GM_5PAIRED)) {
					// skip span ends whose outermost pair cannot form
					// (match_wchlx returns 0 at once, src/find_motif.c:1010-1021)
					const unsigned row = S.duplex >> (bcode_of(L.sq[z]) * 5);
					while (sd >= lsd && !((row >> bcode_of(L.sq[sd])) & 1u))
						sd--;
					if (sd < lsd) {
						L_FR(L, s, 0) = pk16(sd, lsd);
						ph = PH_RET;
						break;
					}
				}
				L_FR(L, s, 0) = pk16(sd, lsd);
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
						L_FR(L, s, 3) = pk16(0, PH_SS_RESUME);
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
					L_FR(L, s, 1) = pk16(z, sd);
					L_FR(L, s, 2) = pk16(s3lim, 0);
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
					L_FR(L, s, 3) = pk16(0, PH_PH_RESUME);
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
					L_FR(L, s, 4) = pk16(sd - e1.minilen - hlen + 1, hlen);
					ph = PH_TR_S;
					break;
				}
				}
				break;
			}

			case PH_WX_BEGIN: {
				// the empty-helix candidate of match_wchlx, src/find_motif.c:986-1006
				const DevSearch &S = sm_ds[s];
				// gm_plan_check refuses seq= on a minlen=0 helix, so the
				// candidate is unconditional
				const uint32_t w1 = L_FR(L, s, 1);
				const int s5 = lo16(w1), s3 = hi16(w1);
				L_FR(L, s, 3) = pk16(FR3_LO(0, 1, 0), PH_WX_RESUME);
				// after this candidate the first pair is tested: hl stays 0
				if (S.kind == K_WC) {
					const int i_len = s3 - s5 + 1;
					if (i_len > S.maxilen) {
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
					L_FR(L, s, 4) = pk16(s5 + S.minilen - 1, 0);
					ph = PH_QU_S1;
				} else {
					// K_PK with minlen 0 is refused by gm_plan_check
					ph = PH_WX_FIRST;
				}
				break;
			}

			case PH_WX_RESUME: {
				const DevSearch &S = sm_ds[s];
				unmark(L, S.d);
				unmark(L, S.d3);
				if (hi16(L_FR(L, s, 2)) == 0) {
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
				const uint32_t w1 = L_FR(L, s, 1), w2 = L_FR(L, s, 2);
				const int s5 = lo16(w1), s3 = hi16(w1), s3lim = lo16(w2);
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
					const int f3 = lo16(L_FR(L, s, 3));
					hl = hi16(w2);
					mpr = f3 & 0xff;
					lbpr = (f3 >> 8) & 1;
					chk = (f3 >> 9) & 1;
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
				L_FR(L, s, 2) = pk16(s3lim, hl);
				L_FR(L, s, 3) = pk16(FR3_LO(mpr, lbpr, 0), PH_WX_RESUME);
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
					L_FR(L, s, 4) = pk16(s5 + hl + S.minilen - 1, 0);
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
				L_FR(L, s, 1) = pk16(s5, f_s3 + 1);
				L_FR(L, s, 5) = pk16(l_s3, i_minl);
				ph = PH_PK_S3;
				break;
			}

			case PH_PK_S3: {
				// find_pknot3 loop over the 3' end, src/find_motif.c:600-606
				const DevSearch &S = sm_ds[s];
				const uint32_t w1 = L_FR(L, s, 1), w5 = L_FR(L, s, 5);
				const int s5 = lo16(w1), s3 = hi16(w1) - 1;
				const int l_s3 = lo16(w5), i_minl = hi16(w5);
				if (s3 < l_s3) {
					ph = PH_PK_S5;
					break;
				}
				L_FR(L, s, 1) = pk16(s5, s3);
				int t3 = s3 - s5 + 1;
				t3 = (t3 - i_minl) / 2;
				t3 = min(t3, S.maxlen);
				L_FR(L, s, 2) = pk16(s3 - t3 + 1, 0);
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
				const uint32_t w4 = L_FR(L, s, 4);
				const int sp = lo16(w4) - 1, hlen = hi16(w4);
				const int z = lo16(L_ZD(L, s)), sd = lo16(L_FR(L, s, 0));
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
				L_FR(L, s, 3) = pk16(lo16(L_FR(L, s, 3)), PH_TR_RESUME);
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
				const int s3 = hi16(L_FR(L, s, 1)), hl = hi16(L_FR(L, s, 2));
				const int s1 = lo16(L_FR(L, s, 4)) + 1;
				if (s1 > s3 - 3 * hl - i3_minl - i2_minl) {
					// this q1/q4 helix is done: back to the extension
					unmark(L, S.d);
					unmark(L, S.d3);
					ph = hl == 0 ? PH_WX_FIRST : PH_WX_EXT;
					break;
				}
				L_FR(L, s, 4) = pk16(s1, s3 - hl - i3_minl + 1);
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
				const uint32_t w4 = L_FR(L, s, 4);
				const int s1 = lo16(w4), s2 = hi16(w4) - 1;
				const int z = lo16(L_FR(L, s, 1)), s3 = hi16(L_FR(L, s, 1)), hl = hi16(L_FR(L, s, 2));
				if (s2 < s1 + 2 * hl + e1.minilen) {
					ph = PH_QU_S1;
					break;
				}
				L_FR(L, s, 4) = pk16(s1, s2);
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
				L_FR(L, s, 3) = pk16(lo16(L_FR(L, s, 3)), PH_QU_RESUME);
				s++;
				ph = PH_ENTER;
				break;
			}

			case PH_RET:
				if (s == 0)
					ph = PH_IDLE;
				else {
					s--;
					ph = hi16(L_FR(L, s, 3));
				}
				break;
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
