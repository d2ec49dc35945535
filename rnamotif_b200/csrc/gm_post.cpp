// gm_post.cpp -- host-only post-filters of libgpumotif over hit records
// (SURVEY section 8 f4): gm_prune_hits = the reference's rmprune, gm_order_hits =
// the order rmfmt prints in.  Plain C++ (no CUDA): also linked into the CPU
// checker of the host driver (oracle/Makefile: rnamotif_hostcheck).
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "gpumotif.h"

extern "C" int gm_post_fail(const char *msg); // sets gm_last_error(), returns -1

static int fail(const char *fmt, ...)
{
	char buf[512];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof buf, fmt, ap);
	va_end(ap);
	return gm_post_fail(buf);
}

// ------------------------------------------------------------------ rmprune
// src/rmprune.c on hit records; names follow the reference (B_SAME ...).
namespace {
enum { PB_UNDEF = -1, PB_SAME = 0, PB_LEFT = 1, PB_DOWN = 2, PB_DIFF = 3 };
struct PruneDetail { int start, stop; };   // DETAIL_T, :68-71
struct PruneHit {
	int comp, start, stop;                 // BLOCK_T::b_comp, b_start, b_stop (:73-83)
	std::vector<PruneDetail> det;
};

// wchlxrel, src/rmprune.c:700-741
int prune_wchlxrel(int comp, const PruneDetail &dp, const PruneDetail &dp_2, const PruneDetail &dp1,
	const PruneDetail &dp1_2)
{
	int lod, rod, lid, rid;
	if (!comp) {
		lod = dp1.start - dp.start;
		rod = dp_2.stop - dp1_2.stop;
		if (lod != rod)
			return PB_DIFF;
		lid = dp.stop - dp1.stop;
		rid = dp1_2.start - dp_2.start;
		if (lid != rid)
			return PB_DIFF;
	} else {
		lod = dp.start - dp1.start;
		rod = dp1_2.stop - dp_2.stop;
		if (lod != rod)
			return PB_DIFF;
		lid = dp1.stop - dp.stop;
		rid = dp_2.start - dp1_2.start;
		if (lid != rid)
			return PB_DIFF;
	}
	if (lod > 0)
		return lid < 0 ? PB_DIFF : PB_DOWN;
	if (lod == 0)
		return lid < 0 ? PB_LEFT : lid == 0 ? PB_SAME : PB_DOWN;
	return lid < 0 ? PB_DIFF : PB_LEFT;
}

// chkrel, src/rmprune.c:649-698
int prune_chkrel(const gm_plan_t *pl, const PruneHit &b, const PruneHit &b1)
{
	int brel = PB_UNDEF, brel1 = PB_UNDEF;
	for (int d = 0; d < pl->n_descr; d++) {
		const gm_elem_t &e = pl->elems[d];
		switch (e.type) {
		case GM_H5:
			brel1 = prune_wchlxrel(b.comp, b.det[d], b.det[e.mates[0]], b1.det[d], b1.det[e.mates[0]]);
			break;
		case GM_P5:
		case GM_T1:
		case GM_Q1: {
			// otherhlxrel, :743-760: every strand at the same place, or different
			brel1 = PB_SAME;
			for (int k = -1; k < e.n_mates && k < 3; k++) {
				const int x = k < 0 ? d : e.mates[k];
				if (b.det[x].start != b1.det[x].start || b.det[x].stop != b1.det[x].stop) {
					brel1 = PB_DIFF;
					break;
				}
			}
			break;
		}
		default:
			brel1 = PB_SAME;
			break;
		}
		if (brel1 == PB_DIFF)
			return PB_DIFF;
		else if (brel == PB_UNDEF)
			brel = brel1;
		else if (brel == PB_SAME)
			brel = brel1;
		else if (brel == PB_DOWN) {
			if (brel1 == PB_LEFT)
				return PB_DIFF;
		} else if (brel == PB_LEFT) {
			if (brel1 == PB_DOWN)
				return PB_DIFF;
		}
	}
	return brel;
}

// rezip_group, src/rmprune.c:400-441, over block[lo, lo + n_group)
void prune_rezip(const gm_plan_t *pl, const std::vector<PruneHit> &blk, size_t lo, size_t n_group, uint8_t *keep)
{
	if (n_group < 2)
		return;
	for (size_t b = n_group - 1; b > 0; b--) {
		if (!keep[lo + b])
			continue;
		for (size_t b1 = b; b1-- > 0;) {
			if (!keep[lo + b1])
				continue;
			const int brel = prune_chkrel(pl, blk[lo + b], blk[lo + b1]);
			if (brel == PB_DOWN)
				keep[lo + b1] = 0;
			else if (brel == PB_LEFT) {
				keep[lo + b] = 0;
				break;
			}
		}
	}
}

// prune_block, src/rmprune.c:332-398
void prune_block(const gm_plan_t *pl, const std::vector<PruneHit> &blk, uint8_t *keep)
{
	const size_t n = blk.size();
	if (n < 2)
		return;
	size_t f_comp = 0;
	while (f_comp < n && !blk[f_comp].comp)
		f_comp++;
	int start = blk[0].start, stop = blk[0].stop;
	size_t lb = 0, b = 0;
	for (; b < f_comp; b++) {
		if (blk[b].start < start || blk[b].stop > stop) {
			prune_rezip(pl, blk, lb, b - lb, keep);
			start = blk[b].start;
			stop = blk[b].stop;
			lb = b;
		}
	}
	prune_rezip(pl, blk, lb, b - lb, keep);
	if (f_comp < n) {
		start = blk[f_comp].start;
		stop = blk[f_comp].stop;
	}
	for (lb = b = f_comp; b < n; b++) {
		if (blk[b].start > start || blk[b].stop < stop) {
			prune_rezip(pl, blk, lb, b - lb, keep);
			start = blk[b].start;
			stop = blk[b].stop;
			lb = b;
		}
	}
	prune_rezip(pl, blk, lb, b - lb, keep);
}
} // namespace

extern "C" int gm_prune_hits(const gm_plan_t *plan, const void *hits, size_t n, size_t stride,
	const int32_t *group, uint8_t *keep)
{
	if (plan == NULL || keep == NULL || (n > 0 && hits == NULL))
		return fail("bad argument");
	if (plan->magic != GM_PLAN_MAGIC || plan->version != GM_PLAN_VERSION)
		return fail("not a plan of this library version");
	const int ND = plan->n_descr;
	if (ND < 1 || ND > GM_MAX_DESCR || stride < sizeof(gm_hit_hdr_t) + (size_t)ND * sizeof(gm_hit_el_t))
		return fail("hit stride %zu too small for %d elements", stride, ND);
	for (int d = 0; d < ND; d++) {
		const gm_elem_t &e = plan->elems[d];
		if ((e.type == GM_H5 || e.type == GM_P5 || e.type == GM_T1 || e.type == GM_Q1) &&
		    (e.n_mates < 1 || e.mates[0] < 0 || e.mates[0] >= ND))
			return fail("element %d: helix without a mate", d);
	}
	const uint8_t *base = static_cast<const uint8_t *>(hits);
	std::vector<PruneHit> blk;
	size_t blk_lo = 0;
	auto flush = [&](size_t hi) {
		prune_block(plan, blk, keep + blk_lo);
		blk.clear();
		blk_lo = hi;
	};
	for (size_t i = 0; i < n; i++) {
		const gm_hit_hdr_t *h = reinterpret_cast<const gm_hit_hdr_t *>(base + i * stride);
		const gm_hit_el_t *el = reinterpret_cast<const gm_hit_el_t *>(h + 1);
		keep[i] = 1;
		const int32_t g = group ? group[i] : (int32_t)h->rec;
		if (i > 0) {
			const gm_hit_hdr_t *hp = reinterpret_cast<const gm_hit_hdr_t *>(base + (i - 1) * stride);
			const int32_t gp = group ? group[i - 1] : (int32_t)hp->rec;
			// a new locus, or the reference's block array is full (BLOCK_SIZE, :82,183-186)
			if (g != gp || blk.size() >= 1000)
				flush(i);
		}
		// enter_block + getdetails, src/rmprune.c:762-818,619-647: the printed start
		// is 1-based; on the complementary strand it counts down from the record's
		// end (only differences matter, so the record length is left out)
		PruneHit ph;
		ph.comp = h->comp ? 1 : 0;
		int len = 0;
		for (int d = 0; d < ND; d++)
			len += el[d].len;
		ph.start = ph.comp ? -(int)el[0].off : (int)el[0].off + 1;
		ph.stop = ph.comp ? ph.start - len + 1 : ph.start + len - 1;
		ph.det.resize(ND);
		int tlen = 0;
		for (int d = 0; d < ND; d++) {
			const int tl = el[d].len > 0 ? el[d].len : 1; // "." for an empty element
			if (!ph.comp) {
				ph.det[d].start = ph.start + tlen;
				ph.det[d].stop = ph.det[d].start + tl - 1;
			} else {
				ph.det[d].start = ph.start - tlen;
				ph.det[d].stop = ph.det[d].start - tl + 1;
			}
			tlen += tl;
		}
		blk.push_back(std::move(ph));
	}
	flush(n);
	return 0;
}

extern "C" int gm_order_hits(const void *hits, size_t n, size_t stride, int n_descr, const double *score,
	const int32_t *name_rank, const int64_t *rec_off, int n_rec, uint32_t *perm)
{
	if (perm == NULL || name_rank == NULL || rec_off == NULL || (n > 0 && hits == NULL) || n > 0xffffffffull)
		return fail("bad argument");
	if (n_descr < 1 || n_descr > GM_MAX_DESCR ||
	    stride < sizeof(gm_hit_hdr_t) + (size_t)n_descr * sizeof(gm_hit_el_t))
		return fail("hit stride %zu too small for %d elements", stride, n_descr);
	const uint8_t *base = static_cast<const uint8_t *>(hits);
	struct Key { double score; int32_t name; int32_t comp; int64_t pos; int32_t len; };
	std::vector<Key> key(n);
	for (size_t i = 0; i < n; i++) {
		const gm_hit_hdr_t *h = reinterpret_cast<const gm_hit_hdr_t *>(base + i * stride);
		const gm_hit_el_t *el = reinterpret_cast<const gm_hit_el_t *>(h + 1);
		if ((int64_t)h->rec >= n_rec)
			return fail("hit %zu: record %u outside the record table", i, h->rec);
		Key &k = key[i];
		k.score = score ? score[i] : 0.0;
		k.name = name_rank[i];
		k.comp = h->comp ? 1 : 0;
		// print_match, src/find_motif.c:1838-1846
		const int64_t slen = rec_off[h->rec + 1] - rec_off[h->rec];
		k.pos = k.comp ? slen - el[0].off : (int64_t)el[0].off + 1;
		k.len = 0;
		for (int d = 0; d < n_descr; d++)
			k.len += el[d].len;
		perm[i] = (uint32_t)i;
	}
	std::stable_sort(perm, perm + n, [&](uint32_t a, uint32_t b) {
		const Key &x = key[a], &y = key[b];
		if (x.score != y.score)
			return x.score > y.score; // -k 2rn
		if (x.name != y.name)
			return x.name < y.name;
		if (x.comp != y.comp)
			return x.comp < y.comp;
		if (x.pos != y.pos)
			return x.pos < y.pos;
		return x.len < y.len;
	});
	return 0;
}


// ---------------------------------------------------------------- score pre-screen (host)
// The interpreter the device runs at the hit sink (gm_score.h), over one hit record and
// the searched strand as characters: the CPU check of the device's decision.
// (GPUMOTIF_SCORE_DEBUG=1: where the host interpreter gave a candidate up, on stderr)
static void gm_score_trace(int pc, int line)
{
	static const bool on = getenv("GPUMOTIF_SCORE_DEBUG") != NULL;
	if (on)
		fprintf(stderr, "gm_score: kept at pc %d (gm_score.h:%d)\n", pc, line);
}
#define GM_SCORE_TRACE(pc, line) gm_score_trace(pc, line)
#include "gm_score.h"

namespace {
struct HostScoreEnv {
	const gm_plan_t *pl;
	const gm_hit_hdr_t *hdr;
	const gm_hit_el_t *els;
	const char *sbuf;
	int sl;
	int ch(int pos) const
	{
		if (pos < 0 || pos >= sl)
			return -1;
		const int c = (unsigned char)sbuf[pos];
		// the device knows a character through its IUPAC code: any other letter is unknown to it
		return strchr("acmgrsvtwyhkdbn", c) != NULL && c != 0 ? c : -1;
	}
	int off(int d) const { return els[d].off; }
	int len(int d) const { return els[d].len; }
	int mpr(int d) const { return els[d].n_mispairs; }
	int mm(int d) const { return els[d].n_mismatches; }
	int comp() const { return hdr->comp; }
	int pos() const { return hdr->comp ? sl - els[0].off : els[0].off + 1; }
	int mlen() const
	{
		int n = 0;
		for (int d = 0; d < pl->n_descr; d++)
			n += els[d].len;
		return n;
	}
	int slen() const { return sl; }
	const gm_elem_t &elem(int d) const { return pl->elems[d]; }
	const gm_pairset_t &pairset(int i) const { return pl->pairsets[i]; }
};
} // namespace

extern "C" int gm_score_prescreen(const gm_plan_t *plan, const gm_score_t *score, const void *hit, const char *sbuf, int slen)
{
	if (plan == NULL || score == NULL || hit == NULL || sbuf == NULL)
		return 0;
	const gm_hit_hdr_t *hdr = static_cast<const gm_hit_hdr_t *>(hit);
	HostScoreEnv env = {plan, hdr, reinterpret_cast<const gm_hit_el_t *>(hdr + 1), sbuf, slen};
	return gm::score_eval(*score, env) == gm::SC_REJECT ? 1 : 0;
}

// ---------------------------------------------------------------- rmfmt [-l | -la]
// src/rmfmt.c:52-376 (the listing form; not -a, the alignment form) over rnamotif's
// output text: "#RM" lines pass through, hit lines are sorted (score descending, name,
// strand, position, length -- the keys rmfmt hands to sort(1), :240-262, compared the
// way sort does in the C locale, whole lines as the last resort) and printed in
// columns as wide as their widest entry; entries longer than 20 characters are
// abbreviated "abc...(n)...xyz" (fcmprs, :474-485); definition lines are dropped.
#include <string>

namespace {
struct FmtLine {
	std::string name;
	std::vector<std::string> sf; // score fields
	std::vector<std::string> of; // strand, position, length, elements
	std::string raw;             // the line as rmfmt writes it to its sort file
};

std::vector<std::string> fmt_split(const std::string &l)
{
	std::vector<std::string> f;
	size_t i = 0;
	while (i < l.size()) {
		while (i < l.size() && (l[i] == ' ' || l[i] == '\t' || l[i] == '\n'))
			i++;
		size_t j = i;
		while (j < l.size() && !(l[j] == ' ' || l[j] == '\t' || l[j] == '\n'))
			j++;
		if (j > i)
			f.push_back(l.substr(i, j - i));
		i = j;
	}
	return f;
}

bool fmt_is_number(const std::string &s) // is_a_number, :431-463
{
	size_t i = 0;
	int mcnt = 0, ecnt = 0, efmt = 0;
	if (i < s.size() && s[i] == '-')
		i++;
	for (; i < s.size() && isdigit((unsigned char)s[i]); i++)
		mcnt++;
	if (i < s.size() && s[i] == '.')
		i++;
	for (; i < s.size() && isdigit((unsigned char)s[i]); i++)
		mcnt++;
	if (i < s.size() && (s[i] == 'e' || s[i] == 'E')) {
		efmt = 1;
		i++;
		if (i < s.size() && s[i] == '-')
			i++;
		for (; i < s.size() && isdigit((unsigned char)s[i]); i++)
			ecnt++;
	}
	return mcnt != 0 && !(efmt && ecnt == 0) && i == s.size();
}

// what sort -n reads: optional blanks, '-', digits, '.', digits
long double fmt_num(const std::string &s)
{
	size_t i = 0;
	while (i < s.size() && (s[i] == ' ' || s[i] == '\t'))
		i++;
	bool neg = false;
	if (i < s.size() && s[i] == '-') {
		neg = true;
		i++;
	}
	long double v = 0, scale = 1;
	for (; i < s.size() && isdigit((unsigned char)s[i]); i++)
		v = v * 10 + (s[i] - '0');
	if (i < s.size() && s[i] == '.')
		for (i++; i < s.size() && isdigit((unsigned char)s[i]); i++) {
			scale /= 10;
			v += (s[i] - '0') * scale;
		}
	return neg ? -v : v;
}
} // namespace

extern "C" int gm_rmfmt(const char *text, size_t n_bytes, int lopt, FILE *out)
{
	if ((text == NULL && n_bytes > 0) || out == NULL || lopt < 0 || lopt > 2)
		return gm_post_fail("gm_rmfmt: bad argument");
	const int DFIELD1 = 4, OFIELD1 = 1, MAXW = 20;
	std::vector<int> ofmt; // 0 left, 1 right
	int n_ofields = 0;
	size_t pos = 0;
	auto next_line = [&](std::string &l) -> bool {
		if (pos >= n_bytes)
			return false;
		const char *e = static_cast<const char *>(memchr(text + pos, '\n', n_bytes - pos));
		const size_t end = e ? (size_t)(e - text) + 1 : n_bytes;
		l.assign(text + pos, end - pos);
		pos = end;
		return true;
	};
	std::string line;
	// header: up to the first definition line (:121-143)
	for (size_t save = pos; next_line(line); save = pos) {
		if (line[0] == '>') {
			pos = save;
			break;
		}
		const std::vector<std::string> f = fmt_split(line);
		if (f.empty() || f[0] != "#RM")
			continue;
		if (f.size() < 2)
			continue;
		if (f[1] == "descr") {
			// getfmt, :404-428
			n_ofields = (int)f.size() - 2 + DFIELD1;
			ofmt.assign(n_ofields, 0);
			for (int k = 1; k < DFIELD1; k++)
				ofmt[k] = 1;
			for (size_t k = 2; k < f.size(); k++) {
				const std::string t = f[k].substr(0, 2);
				ofmt[k - 2 + DFIELD1] = (t == "h3" || t == "t2" || t == "q2" || t == "q4") ? 1 : 0;
			}
		}
		fputs(line.c_str(), out);
	}
	std::vector<int> omaxw(std::max(n_ofields, 1), 0);
	std::vector<FmtLine> rows;
	int smaxw = 0, m_sfields = 0, n_sfields1 = 0, scored = 1, sorted = 1, n_sfields = 0;
	size_t raw_bytes = 0;
	while (next_line(line)) {
		if (line[0] == '#' || line[0] == '>')
			continue;
		std::vector<std::string> f = fmt_split(line);
		const int n_fields = (int)f.size();
		const int df1 = n_fields - (n_ofields - DFIELD1), of1 = df1 - (DFIELD1 - OFIELD1);
		n_sfields = df1 - DFIELD1;
		if (n_fields == 0 || of1 < 1 || n_sfields < 0)
			return gm_post_fail("gm_rmfmt: a hit line does not fit the #RM descr line");
		m_sfields = std::max(m_sfields, n_sfields);
		if (n_sfields > 1)
			scored = 0;
		if (n_sfields1 == 0)
			n_sfields1 = n_sfields;
		else if (n_sfields != m_sfields)
			sorted = 0;
		FmtLine r;
		r.name = f[0];
		for (int k = 0; k < n_sfields; k++)
			r.sf.push_back(f[k + 1]);
		for (int k = of1; k < n_fields; k++)
			r.of.push_back(f[k]);
		if (scored)
			scored = !r.sf.empty() && fmt_is_number(r.sf[0]);
		if (lopt) {
			// -l: the locus name is what follows the last '|' (or lies between the last two
			// when nothing follows); -la: the accession between the last two (:181-212)
			size_t vb = std::string::npos, lvb = std::string::npos;
			for (size_t k = 0; k < r.name.size(); k++)
				if (r.name[k] == '|') {
					lvb = vb;
					vb = k;
				}
			if (vb != std::string::npos) {
				if (lopt == 1) {
					if (vb + 1 < r.name.size())
						r.name = r.name.substr(vb + 1);
					else if (lvb != std::string::npos)
						r.name = r.name.substr(lvb + 1, vb - lvb - 1);
					else
						r.name = r.name.substr(0, vb);
				} else if (lvb != std::string::npos)
					r.name = r.name.substr(lvb + 1, vb - lvb - 1);
			}
		}
		omaxw[0] = std::max(omaxw[0], (int)r.name.size());
		int fw = 0;
		for (const std::string &x : r.sf)
			fw += (int)x.size();
		fw += n_sfields - 1;
		smaxw = std::max(smaxw, fw);
		for (int k = OFIELD1; k < n_ofields && k - 1 < (int)r.of.size(); k++) {
			std::string &x = r.of[k - 1];
			if (k >= DFIELD1 && (int)x.size() > MAXW) {
				char tmp[128];
				snprintf(tmp, sizeof tmp, "%.3s...(%d)...%.3s", x.c_str(), (int)x.size(), x.c_str() + x.size() - 3);
				x = tmp;
			}
			omaxw[k] = std::max(omaxw[k], (int)x.size());
		}
		r.raw = r.name;
		for (const std::string &x : r.sf)
			r.raw += " " + x;
		for (const std::string &x : r.of)
			r.raw += " " + x;
		raw_bytes += r.raw.size() + 1;
		rows.push_back(std::move(r));
	}
	if (raw_bytes > 30000000) // SMAX: files larger than this are not sorted (:216-217)
		scored = sorted = 0;
	if (scored || sorted) {
		const int ns = n_sfields; // (of the last line, as rmfmt's command line has it)
		auto key = [&](const FmtLine &r, int k) -> long double { // field k (1-based) of the raw line, numeric
			if (k == 1)
				return fmt_num(r.name);
			if (k - 2 < (int)r.sf.size())
				return fmt_num(r.sf[k - 2]);
			const int o = k - 2 - (int)r.sf.size();
			return o < (int)r.of.size() ? fmt_num(r.of[o]) : 0;
		};
		std::stable_sort(rows.begin(), rows.end(), [&](const FmtLine &a, const FmtLine &b) {
			if (scored) {
				const long double x = key(a, 2), y = key(b, 2);
				if (x != y)
					return x > y; // -k 2rn,2
			}
			const int c = a.name.compare(b.name); // -k 1,1
			if (c != 0)
				return c < 0;
			const int k0 = scored ? 3 : ns + 2;
			for (int k = k0; k < k0 + 3; k++) {
				const long double x = key(a, k), y = key(b, k);
				if (x != y)
					return x < y;
			}
			return a.raw < b.raw; // sort's last resort: the whole line, bytewise in the C locale
		});
	}
	for (const FmtLine &r : rows) {
		auto put = [&](int k, const std::string &x) {
			const int w = (int)x.size();
			if (ofmt[k] == 0) {
				if (k != 0)
					fputc(' ', out);
				fputs(x.c_str(), out);
				for (int s = 0; s < omaxw[k] - w; s++)
					fputc(' ', out);
			} else {
				fputc(' ', out);
				for (int s = 0; s < omaxw[k] - w; s++)
					fputc(' ', out);
				fputs(x.c_str(), out);
			}
		};
		put(0, r.name);
		fputc(' ', out);
		if (m_sfields == 1) {
			const std::string &x = r.sf.empty() ? std::string() : r.sf[0];
			for (int s = 0; s < smaxw - (int)x.size(); s++)
				fputc(' ', out);
			fputs(x.c_str(), out);
		} else {
			int fs = 0;
			for (size_t k = 0; k < r.sf.size(); k++) {
				if (k != 0) {
					fputc(' ', out);
					fs++;
				}
				fputs(r.sf[k].c_str(), out);
				fs += (int)r.sf[k].size();
			}
			for (; fs < smaxw; fs++)
				fputc(' ', out);
		}
		for (int k = OFIELD1; k < n_ofields && k - 1 < (int)r.of.size(); k++)
			put(k, r.of[k - 1]);
		fputc('\n', out);
	}
	return 0;
}
