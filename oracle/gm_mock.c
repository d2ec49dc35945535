/*
 * oracle/gm_mock.c -- TEST INFRASTRUCTURE.  A stand-in for the part of the C ABI
 * of libgpumotif (include/gpumotif.h) that the reference-side host driver
 * (rnamotif_b200/host/rm_gpu_main.c) calls, with the candidates coming from the
 * plain-C oracle port (ref_search.c) instead of the device.  Linked ONLY into
 * oracle/_ref/rnamotif_hostcheck, which tests/test_host_driver_cpu.py runs to
 * check the HOST logic of the driver -- batching of records, sharding of a batch
 * into start ranges and the ordered merge, the replay of the sink's tail through
 * the reference's score program and printer -- byte for byte against the
 * reference's stdout on a machine without a GPU.  Never part of the product:
 * libgpumotif.so has no CPU search path.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "gpumotif.h"

typedef struct gmo_stats {
	uint64_t n_starts, n_pair_evals, n_chk_seq, n_regex_steps, n_candidates;
} gmo_stats_t;
extern int64_t gmo_scan_db(const gm_plan_t *pl, const char *seq, const int64_t *rec_off, int n_rec, int both,
	void *out, size_t out_cap, gmo_stats_t *stats);
extern size_t gmo_hit_stride(const gm_plan_t *pl);

struct gm_ctx {
	gm_plan_t plan;
	char *seq;
	int64_t *off;
	int64_t *hdr;   /* gm_db_upload_fastn: text offset of every record's '>' */
	char *wins;     /* gm_hit_windows */
	int n_rec;
	int64_t lo, hi;
	int strands;
	char *hits;
	size_t n, stride;
	gm_score_t *score;  /* gm_ctx_set_score: candidates it rejects are dropped like the device does */
	uint64_t n_rejected;
};

static char mock_err[512] = "";
const char *gm_last_error(void) { return mock_err; }
/* error sink of gm_post.cpp (gm_prune_hits / gm_order_hits, linked as they are) */
int gm_post_fail(const char *msg)
{
	snprintf(mock_err, sizeof mock_err, "%s", msg);
	return -1;
}

int gm_ctx_create(gm_ctx **out, const gm_plan_t *plan, int device)
{
	gm_ctx *c = calloc(1, sizeof *c);
	(void)device;
	if (c == NULL)
		return -1;
	c->plan = *plan;
	c->stride = gmo_hit_stride(plan);
	*out = c;
	return 0;
}

void gm_ctx_destroy(gm_ctx *c)
{
	if (c == NULL)
		return;
	free(c->seq);
	free(c->off);
	free(c->hdr);
	free(c->wins);
	free(c->hits);
	free(c->score);
	free(c);
}

int gm_ctx_set_score(gm_ctx *c, const gm_score_t *score)
{
	free(c->score);
	c->score = NULL;
	if (score != NULL && score->present) {
		c->score = malloc(sizeof *score);
		if (c->score == NULL)
			return -1;
		*c->score = *score;
	}
	return 0;
}

int gm_db_upload_chars(gm_ctx *c, const char *seq, const int64_t *rec_off, int n_rec)
{
	const int64_t total = rec_off[n_rec];
	free(c->seq);
	free(c->off);
	c->seq = malloc((size_t)total + 1);
	c->off = malloc((size_t)(n_rec + 1) * sizeof *c->off);
	if (c->seq == NULL || c->off == NULL)
		return -1;
	memcpy(c->seq, seq, (size_t)total);
	memcpy(c->off, rec_off, (size_t)(n_rec + 1) * sizeof *c->off);
	c->n_rec = n_rec;
	return 0;
}

/* FN_fgetseq (src/dbutil.c:42-128) as the two-state machine gm_fastn.cuh describes:
 * outside a header every '>' starts a record and every isalpha character is kept;
 * a header runs to its newline */
int gm_db_upload_fastn(gm_ctx *c, const char *text, size_t n_bytes)
{
	size_t i, n_rec = 0, n_ch = 0, cap_rec = 1024;
	int in_hdr = 0;
	if (n_bytes > 0 && text[0] != '>') {
		snprintf(mock_err, sizeof mock_err, "fastn text does not begin with '>'");
		return -1;
	}
	free(c->seq);
	free(c->off);
	free(c->hdr);
	c->seq = malloc(n_bytes + 1);
	c->off = malloc((cap_rec + 1) * sizeof *c->off);
	c->hdr = malloc((cap_rec + 1) * sizeof *c->hdr);
	if (c->seq == NULL || c->off == NULL || c->hdr == NULL)
		return -1;
	for (i = 0; i < n_bytes; i++) {
		const unsigned char ch = (unsigned char)text[i];
		if (in_hdr) {
			if (ch == '\n')
				in_hdr = 0;
		} else if (ch == '>') {
			if (n_rec == cap_rec) {
				cap_rec *= 2;
				c->off = realloc(c->off, (cap_rec + 1) * sizeof *c->off);
				c->hdr = realloc(c->hdr, (cap_rec + 1) * sizeof *c->hdr);
				if (c->off == NULL || c->hdr == NULL)
					return -1;
			}
			c->off[n_rec] = (int64_t)n_ch;
			c->hdr[n_rec] = (int64_t)i;
			n_rec++;
			in_hdr = 1;
		} else if ((ch >= 'a' && ch <= 'z') || (ch >= 'A' && ch <= 'Z'))
			c->seq[n_ch++] = (char)ch;
	}
	c->off[n_rec] = (int64_t)n_ch;
	c->hdr[n_rec] = (int64_t)n_bytes;
	c->n_rec = (int)n_rec;
	return 0;
}

int gm_db_records(const gm_ctx *c, const int64_t **rec_off, const int64_t **hdr_off, int *n_rec)
{
	if (rec_off)
		*rec_off = c->off;
	if (hdr_off)
		*hdr_off = c->hdr;
	if (n_rec)
		*n_rec = c->n_rec;
	return 0;
}

int64_t gm_db_total_nt(const gm_ctx *c) { return c->n_rec > 0 || c->off ? c->off[c->n_rec] : 0; }

int gm_host_alloc(void **out, size_t n_bytes)
{
	*out = malloc(n_bytes ? n_bytes : 1);
	return *out ? 0 : -1;
}
void gm_host_free(void *p) { free(p); }

int gm_scan_launch(gm_ctx *c, int64_t g_begin, int64_t g_end, int strands)
{
	c->lo = g_begin;
	c->hi = g_end;
	c->strands = strands;
	return 0;
}

/* the oracle scans whole records; a context owns the starts whose 5' end (in the
 * searched strand) falls on a nucleotide in [lo, hi), like gm_scan */
int gm_scan_finish(gm_ctx *c)
{
	size_t cap = 1 << 16, i, k = 0;
	int64_t n;
	c->n_rejected = 0;
	gmo_stats_t st;
	for (;;) {
		free(c->hits);
		c->hits = malloc(cap * c->stride);
		if (c->hits == NULL)
			return -1;
		memset(&st, 0, sizeof st);
		n = gmo_scan_db(&c->plan, c->seq, c->off, c->n_rec, c->strands == 2, c->hits, cap * c->stride, &st);
		if (n < 0) {
			snprintf(mock_err, sizeof mock_err, "oracle scan failed");
			return -1;
		}
		if ((size_t)n <= cap)
			break;
		cap = (size_t)n;
	}
	for (i = 0; i < (size_t)n; i++) {
		const gm_hit_hdr_t *h = (const gm_hit_hdr_t *)(c->hits + i * c->stride);
		const int64_t slen = c->off[h->rec + 1] - c->off[h->rec];
		const int64_t g = c->off[h->rec] + (h->comp ? slen - 1 - (int64_t)h->szero : (int64_t)h->szero);
		if (g >= c->lo && g < c->hi && c->score != NULL) {
			/* the searched strand of this record as fm_sbuf holds it */
			char *sb = malloc((size_t)slen + 1);
			int64_t p;
			int rej;
			if (sb == NULL)
				return -1;
			for (p = 0; p < slen; p++) {
				int ch = (unsigned char)c->seq[c->off[h->rec] + (h->comp ? slen - 1 - p : p)] | 0x20;
				if (ch == 'u')
					ch = 't';
				if (h->comp)
					ch = ch == 'a' ? 't' : ch == 'c' ? 'g' : ch == 'g' ? 'c' : ch == 't' ? 'a' : 'n';
				sb[p] = (char)ch;
			}
			sb[slen] = 0;
			rej = gm_score_prescreen(&c->plan, c->score, h, sb, (int)slen);
			free(sb);
			if (rej) {
				c->n_rejected++;
				continue;
			}
		}
		if (g >= c->lo && g < c->hi) {
			if (k != i)
				memmove(c->hits + k * c->stride, c->hits + i * c->stride, c->stride);
			k++;
		}
	}
	c->n = k;
	return 0;
}

int gm_scan(gm_ctx *c, int64_t g_begin, int64_t g_end, int strands)
{
	return gm_scan_launch(c, g_begin, g_end, strands) || gm_scan_finish(c);
}

/* what fm_sbuf holds around every candidate: lower case, u -> t; on the complementary
 * strand mk_rcmp's letters (src/rnamot.c:193-216); 0 outside the record */
int gm_hit_windows(gm_ctx *c, int lead, int trail, const char **win, size_t *stride)
{
	const int W = c->plan.dmaxlen < c->plan.windowsize ? c->plan.dmaxlen : c->plan.windowsize;
	const size_t wlen = (size_t)((lead + W + trail + 1 + 7) & ~7);
	size_t i, j;
	free(c->wins);
	c->wins = calloc(c->n ? c->n : 1, wlen);
	if (c->wins == NULL)
		return -1;
	for (i = 0; i < c->n; i++) {
		const gm_hit_hdr_t *h = (const gm_hit_hdr_t *)(c->hits + i * c->stride);
		const int64_t slen = c->off[h->rec + 1] - c->off[h->rec];
		const char *rec = c->seq + c->off[h->rec];
		for (j = 0; j < wlen; j++) {
			const int64_t p = (int64_t)h->szero - lead + (int64_t)j;
			int ch;
			if (p < 0 || p >= slen)
				continue;
			ch = (unsigned char)rec[h->comp ? slen - 1 - p : p] | 0x20;
			if (ch == 'u')
				ch = 't';
			if (h->comp)
				ch = ch == 'a' ? 't' : ch == 'c' ? 'g' : ch == 'g' ? 'c' : ch == 't' ? 'a' : 'n';
			c->wins[i * wlen + j] = (char)ch;
		}
	}
	*win = c->wins;
	*stride = wlen;
	return 0;
}

int gm_hits(const gm_ctx *c, const void **hits, size_t *n, size_t *stride)
{
	*hits = c->hits;
	*n = c->n;
	*stride = c->stride;
	return 0;
}

int gm_stats(const gm_ctx *c, gm_scan_stats_t *out)
{
	memset(out, 0, sizeof *out);
	out->n_hits = c->n;
	out->n_score_rejected = c->n_rejected;
	return 0;
}
