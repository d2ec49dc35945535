/*
 * rm_gpu_main.c -- rnamotif with the search on the GPU.  Same command line,
 * same stdout.  The front end, the score program, the energy code and the
 * printer are the reference's (linked from its own objects); the record loop
 * of src/rnamot.c:125-190 is re-arranged into batches:
 *
 *     read a batch of records with the reference's fgetseq
 *     gm_db_upload_chars + gm_scan          (libgpumotif, include/gpumotif.h)
 *     for every candidate, in enumeration order: GM_replay_hit (rm_replay.c)
 *
 * Environment: GPUMOTIF_DEVICES ("0,1,..": the batch is cut into that many
 * ranges of start positions, one GPU each, no exchange; default GPUMOTIF_DEVICE
 * or 0), GPUMOTIF_BATCH_NT (default 64 M), GPUMOTIF_STATS=1 prints per-batch
 * timings to stderr.  This in-process sharding stands in for the MPI file farm of
 * src/mrnamotif.c:105-192.
 * GPUMOTIF_PRUNE=1: the output of `rnamotif | rmprune` in one pass -- the hits the
 * score program accepts are captured instead of printed, gm_prune_hits() takes
 * rmprune's decision on their records (src/rmprune.c), and only the kept ones are
 * written.  (Hits a score program HOLDs and RELEASEs are printed by score.c itself
 * and pass through unpruned.)
 */
#define _GNU_SOURCE /* open_memstream */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "log.h"
#include "rmdefs.h"
#include "rnamot.h"
#include "dbutil.h"
#include "gpumotif.h"

extern int rm_error;
extern ARGS_T *rm_args;
extern FILE *rm_dbfp;

extern int gm_rm_compile(int, char *[]);
extern int gm_flatten_plan(gm_plan_t *, char *, size_t);
extern void GM_replay_strand(char[], char[], int, int, char[]);
extern int GM_replay_hit(const gm_hit_hdr_t *, const gm_hit_el_t *);
extern int GM_replay_hit_to(const gm_hit_hdr_t *, const gm_hit_el_t *, FILE *);

typedef struct {
	char *sid, *sdef;
	int64_t off;
	int slen;
} REC_T;

/* reverse complement as the reference builds it in place (src/rnamot.c:193-216):
 * a<->t c<->g (u as t), anything else -> n */
static void revcomp_into(const char *src, int slen, char *dst)
{
	static char cmp[256];
	static int init;
	int i;
	if (!init) {
		init = 1;
		memset(cmp, 'n', sizeof cmp);
		cmp['a'] = cmp['A'] = 't';
		cmp['c'] = cmp['C'] = 'g';
		cmp['g'] = cmp['G'] = 'c';
		cmp['t'] = cmp['T'] = 'a';
		cmp['u'] = cmp['U'] = 'a';
	}
	for (i = 0; i < slen; i++)
		dst[i] = cmp[(unsigned char)src[slen - 1 - i]];
	dst[slen] = '\0';
}

#define MAX_GPUS 16

/* enumeration order of the reference: record, strand, start, DFS rank */
static size_t hit_stride;
static int hit_cmp(const void *pa, const void *pb)
{
	const gm_hit_hdr_t *a = *(const gm_hit_hdr_t *const *)pa, *b = *(const gm_hit_hdr_t *const *)pb;
	if (a->rec != b->rec)
		return a->rec < b->rec ? -1 : 1;
	if (a->comp != b->comp)
		return a->comp < b->comp ? -1 : 1;
	if (a->szero != b->szero)
		return a->szero < b->szero ? -1 : 1;
	return a->seq < b->seq ? -1 : a->seq > b->seq;
}

static void die_gm(const char *what)
{
	fprintf(stderr, "rnamotif_gpu: %s: %s\n", what, gm_last_error());
	exit(1);
}

int main(int argc, char *argv[])
{
	static gm_plan_t plan;
	static char sid[SID_SIZE], sdef[SDEF_SIZE];
	char err[512];
	gm_ctx *ctxs[MAX_GPUS];
	int devices[MAX_GPUS], n_gpus = 0, g;
	const gm_hit_hdr_t **order = NULL;
	size_t order_cap = 0;
	IDENT_T *ip;
	int chk_both_strs, show_progress, ecnt = 0, stats = 0, eof = 0;
	int (*fgetseq)(FILE *, char *, int, char *, int, char *);
	int64_t batch_nt = 64ll << 20, buf_cap, used;
	char *buf, *rcbuf = NULL;
	int rc_cap = 0;
	REC_T *recs = NULL;
	int64_t *offs = NULL;
	int n_recs, cap_recs = 0;
	const char *ev;
	/* GPUMOTIF_PRUNE: accepted hits of the batch, their text and their records */
	int prune = 0;
	FILE *cap = NULL;
	char *cap_buf = NULL;
	size_t cap_len = 0, n_acc = 0, acc_cap = 0;
	size_t *acc_end = NULL;          /* end of hit k's text in the capture */
	char *acc_hits = NULL;           /* their hit records, contiguous */
	int32_t *acc_group = NULL;       /* rmprune's blocks: locus name up to the first '.' */
	uint8_t *acc_keep = NULL;

	gm_rm_compile(argc, argv);
	if (gm_flatten_plan(&plan, err, sizeof err)) {
		fprintf(stderr, "rnamotif_gpu: descriptor cannot run on the device: %s\n", err);
		exit(1);
	}
	if ((ev = getenv("GPUMOTIF_DEVICES")) != NULL && *ev) {
		char *copy = strdup(ev), *tok;
		for (tok = strtok(copy, ","); tok != NULL && n_gpus < MAX_GPUS; tok = strtok(NULL, ","))
			devices[n_gpus++] = atoi(tok);
		free(copy);
	}
	if (n_gpus == 0) {
		devices[n_gpus++] = (ev = getenv("GPUMOTIF_DEVICE")) != NULL ? atoi(ev) : 0;
	}
	if ((ev = getenv("GPUMOTIF_BATCH_NT")) != NULL && atoll(ev) > 0)
		batch_nt = atoll(ev);
	if ((ev = getenv("GPUMOTIF_STATS")) != NULL)
		stats = atoi(ev);
	if ((ev = getenv("GPUMOTIF_PRUNE")) != NULL)
		prune = atoi(ev);
	for (g = 0; g < n_gpus; g++)
		if (gm_ctx_create(&ctxs[g], &plan, devices[g]))
			die_gm("gm_ctx_create");

	ip = RM_find_id("chk_both_strs");
	chk_both_strs = ip == NULL ? 1 : ip->i_val.v_value.v_ival;
	ip = RM_find_id("show_progress");
	show_progress = ip == NULL ? 0 : ip->i_val.v_value.v_ival;

	if (rm_args->a_dbfmt == NULL || !strcmp(rm_args->a_dbfmt, DT_FASTN))
		fgetseq = FN_fgetseq;
	else if (!strcmp(rm_args->a_dbfmt, DT_PIR))
		fgetseq = PIR_fgetseq;
	else if (!strcmp(rm_args->a_dbfmt, DT_GENBANK))
		fgetseq = GB_fgetseq;
	else {
		rm_error = TRUE;
		LOG_ERROR("unknown data format %s.", rm_args->a_dbfmt);
		exit(1);
	}
	rm_dbfp = DB_fnext(rm_dbfp, &rm_args->a_c_dbfname, rm_args->a_n_dbfname, rm_args->a_dbfname);
	if (rm_dbfp == NULL)
		exit(1);

	buf_cap = batch_nt + rm_args->a_maxslen + 16;
	buf = malloc((size_t)buf_cap);
	if (buf == NULL) {
		rm_error = TRUE;
		LOG_ERROR("can't allocate sbuf (s_sbuf=%lld)", (long long)buf_cap);
		exit(1);
	}
	if (RM_fm_init())
		exit(1);

	RM_setprog(P_BEGIN);
	RM_score(0, 0, NULL, NULL);
	RM_setprog(P_MAIN);

	while (!eof) {
		size_t n_hits = 0, stride = 0, h;
		int r;

		/* ---- read a batch ---- */
		n_recs = 0;
		used = 0;
		while (used < batch_nt) {
			int slen = fgetseq(rm_dbfp, sid, SDEF_SIZE, sdef, rm_args->a_maxslen, buf + used);
			if (slen == EOF) {
				rm_dbfp = DB_fnext(rm_dbfp, &rm_args->a_c_dbfname, rm_args->a_n_dbfname,
					rm_args->a_dbfname);
				if (rm_dbfp == NULL) {
					eof = 1;
					break;
				}
				continue;
			}
			ecnt++;
			if (show_progress && ecnt % show_progress == 0)
				fprintf(stderr, "%s: %7d: %s\n", argv[0], ecnt, sid);
			if (n_recs == cap_recs) {
				cap_recs = cap_recs ? 2 * cap_recs : 1024;
				recs = realloc(recs, cap_recs * sizeof *recs);
				offs = realloc(offs, (cap_recs + 1) * sizeof *offs);
				if (recs == NULL || offs == NULL) {
					fprintf(stderr, "rnamotif_gpu: out of memory\n");
					exit(1);
				}
			}
			recs[n_recs].sid = strdup(sid);
			recs[n_recs].sdef = strdup(sdef);
			recs[n_recs].off = used;
			recs[n_recs].slen = slen;
			n_recs++;
			used += slen; /* the NUL fgetseq wrote is overwritten by the next record;
				       * the replay restores it per record below */
		}
		if (n_recs == 0)
			break;
		for (r = 0; r < n_recs; r++)
			offs[r] = recs[r].off;
		offs[n_recs] = used;

		/* ---- search on the device(s): GPU g owns the starts in [g, g+1) * used / n_gpus ---- */
		for (g = 0; g < n_gpus; g++)
			if (gm_db_upload_chars(ctxs[g], buf, offs, n_recs))
				die_gm("gm_db_upload_chars");
		for (g = 0; g < n_gpus; g++)
			if (gm_scan_launch(ctxs[g], used * g / n_gpus, used * (g + 1) / n_gpus, chk_both_strs ? 2 : 1))
				die_gm("gm_scan_launch");
		for (g = 0; g < n_gpus; g++) {
			const void *hp;
			size_t n, i;
			if (gm_scan_finish(ctxs[g]))
				die_gm("gm_scan_finish");
			if (gm_hits(ctxs[g], &hp, &n, &stride))
				die_gm("gm_hits");
			if (n_hits + n > order_cap) {
				order_cap = 2 * (n_hits + n) + 1024;
				order = realloc(order, order_cap * sizeof *order);
				if (order == NULL) {
					fprintf(stderr, "rnamotif_gpu: out of memory\n");
					exit(1);
				}
			}
			for (i = 0; i < n; i++)
				order[n_hits + i] = (const gm_hit_hdr_t *)((const char *)hp + i * stride);
			n_hits += n;
			if (stats) {
				gm_scan_stats_t st;
				gm_stats(ctxs[g], &st);
				fprintf(stderr, "rnamotif_gpu: gpu %d batch %d records %lld nt: upload %.2f ms kernel %.2f ms "
					"d2h %.2f ms sort %.2f ms, %llu candidates, %u retries\n", devices[g], n_recs,
					(long long)used, st.h2d_ms, st.kernel_ms, st.d2h_ms, st.sort_ms,
					(unsigned long long)st.n_hits, st.n_retries);
			}
		}
		/* each GPU's list is sorted; with several GPUs merge them into one order
		 * (the score program is stateful: src/score.c HOLD/RELEASE) */
		hit_stride = stride;
		if (n_gpus > 1 && n_hits > 1)
			qsort(order, n_hits, sizeof *order, hit_cmp);

		if (prune) {
			cap = open_memstream(&cap_buf, &cap_len);
			if (cap == NULL) {
				fprintf(stderr, "rnamotif_gpu: open_memstream failed\n");
				exit(1);
			}
			n_acc = 0;
		}
		/* ---- replay the sink's tail in enumeration order ---- */
		for (h = 0; h < n_hits;) {
			const gm_hit_hdr_t *hdr = order[h];
			const uint32_t rec = hdr->rec;
			const int comp = hdr->comp;
			REC_T *rp = &recs[rec];
			char *sb = buf + rp->off, saved;
			saved = sb[rp->slen];
			sb[rp->slen] = '\0';
			if (comp) {
				if (rp->slen + 1 > rc_cap) {
					rc_cap = rp->slen + 1;
					rcbuf = realloc(rcbuf, rc_cap);
					if (rcbuf == NULL) {
						fprintf(stderr, "rnamotif_gpu: out of memory\n");
						exit(1);
					}
				}
				revcomp_into(sb, rp->slen, rcbuf);
				GM_replay_strand(rp->sid, rp->sdef, 1, rp->slen, rcbuf);
			} else
				GM_replay_strand(rp->sid, rp->sdef, 0, rp->slen, sb);
			for (; h < n_hits; h++) {
				hdr = order[h];
				if (hdr->rec != rec || hdr->comp != comp)
					break;
				if (!prune) {
					GM_replay_hit(hdr, (const gm_hit_el_t *)(hdr + 1));
					continue;
				}
				if (GM_replay_hit_to(hdr, (const gm_hit_el_t *)(hdr + 1), cap)) {
					fflush(cap);
					if (n_acc == acc_cap) {
						acc_cap = acc_cap ? 2 * acc_cap : 1024;
						acc_end = realloc(acc_end, acc_cap * sizeof *acc_end);
						acc_hits = realloc(acc_hits, acc_cap * stride);
						acc_group = realloc(acc_group, acc_cap * sizeof *acc_group);
						acc_keep = realloc(acc_keep, acc_cap);
						if (acc_end == NULL || acc_hits == NULL || acc_group == NULL || acc_keep == NULL) {
							fprintf(stderr, "rnamotif_gpu: out of memory\n");
							exit(1);
						}
					}
					acc_end[n_acc] = cap_len;
					memcpy(acc_hits + n_acc * stride, hdr, stride);
					acc_group[n_acc] = (int32_t)rec;
					n_acc++;
				}
			}
			sb[rp->slen] = saved;
		}
		if (prune) {
			size_t k, from = 0;
			int32_t gid = 0;
			fclose(cap);
			/* rmprune's blocks are runs of hits whose locus names agree up to the
			 * first '.' (getname, src/rmprune.c:312-330): number them */
			for (k = 0; k < n_acc; k++) {
				if (k > 0 && acc_group[k] != (int32_t)((const gm_hit_hdr_t *)(acc_hits + (k - 1) * stride))->rec) {
					const char *a = recs[acc_group[k]].sid;
					const char *b = recs[((const gm_hit_hdr_t *)(acc_hits + (k - 1) * stride))->rec].sid;
					size_t la = strcspn(a, "."), lb = strcspn(b, ".");
					if (la != lb || strncmp(a, b, la))
						gid++;
				}
				acc_group[k] = gid; /* (the record is still in the copied hit header) */
			}
			if (gm_prune_hits(&plan, acc_hits, n_acc, stride, acc_group, acc_keep))
				die_gm("gm_prune_hits");
			for (k = 0; k < n_acc; k++) {
				const char *t = cap_buf + from, *e = cap_buf + acc_end[k];
				/* the first hit of the run carries print_match's "#RM" header lines:
				 * rmprune passes them through whatever becomes of the hit */
				while (t < e && *t != '>') {
					const char *nl = memchr(t, '\n', (size_t)(e - t));
					nl = nl ? nl + 1 : e;
					fwrite(t, 1, (size_t)(nl - t), stdout);
					t = nl;
				}
				if (acc_keep[k])
					fwrite(t, 1, (size_t)(e - t), stdout);
				from = acc_end[k];
			}
			free(cap_buf);
			cap_buf = NULL;
			cap_len = 0;
		}
		for (r = 0; r < n_recs; r++) {
			free(recs[r].sid);
			free(recs[r].sdef);
		}
	}

	RM_setprog(P_END);
	RM_score(0, 0, NULL, NULL);
	for (g = 0; g < n_gpus; g++)
		gm_ctx_destroy(ctxs[g]);
	exit(0);
}
