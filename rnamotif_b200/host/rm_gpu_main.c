/*
 * rm_gpu_main.c -- rnamotif with the search on the GPU.  Same command line,
 * same stdout.  The front end, the score program, the energy code and the
 * printer are the reference's (linked from its own objects); the record loop
 * of src/rnamot.c:125-190 is re-arranged into batches that flow through a
 * pipeline:
 *
 *   read     the FASTA text of a batch of whole records, straight from the file
 *            into pinned memory (pread; no character is looked at on the host)
 *   search   gm_db_upload_fastn (the reader of src/dbutil.c:42-128 runs on the
 *            device) + gm_scan + gm_hit_windows        (include/gpumotif.h)
 *   replay   for every candidate, in enumeration order: the tail of the hit sink
 *            (RM_score + print_match, rm_replay.c) over the window of the searched
 *            strand that came back with the candidate
 *
 * Two search workers per GPU, each with its own context, take batches in file
 * order; the replay runs on the main thread in batch order while the workers are
 * busy with the next batches (the score program is stateful and not re-entrant:
 * src/score.c HOLD/RELEASE).  With several GPUs (GPUMOTIF_DEVICES=0,1,..)
 * batches are dealt to the workers of all of them -- whole records only, so no
 * halo and no exchange.  This stands in for the MPI file farm of
 * src/mrnamotif.c:105-192.
 *
 * Whatever the device reader does not do exactly like FN_fgetseq -- PIR / GenBank
 * input, standard input, an entry without a name, a record longer than -maxslen --
 * goes through the host reader (the reference's own fgetseq, one batch at a time:
 * read, gm_db_upload_chars, gm_scan, replay); the pipeline hands over at the batch
 * where it meets such a case, at that batch's file offset.
 *
 * Environment: GPUMOTIF_DEVICES / GPUMOTIF_DEVICE, GPUMOTIF_BATCH_NT (nucleotides
 * or text bytes per batch, default 128 M), GPUMOTIF_READER=host (host reader
 * only), GPUMOTIF_STATS=1 (per-batch timings and a summary on stderr),
 * GPUMOTIF_PRUNE=1: the output of `rnamotif | rmprune` in one pass -- the hits
 * the score program accepts are captured instead of printed, gm_prune_hits() takes
 * rmprune's decision on their records (src/rmprune.c), and only the kept ones are
 * written.  Blocks are cut at batch boundaries; hits a score program HOLDs and
 * RELEASEs are printed by score.c itself and pass through unpruned.
 * GPUMOTIF_FMT=1 | l | la: the output of `rnamotif | rmfmt`, `| rmfmt -l`, `| rmfmt -la`
 * in one program (gm_rmfmt, src/rmfmt.c): stdout is collected in a temporary file
 * and formatted at the end.
 */
#define _GNU_SOURCE /* open_memstream, pread */
#include <ctype.h>
#include <fcntl.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include "log.h"
#include "rmdefs.h"
#include "rnamot.h"
#include "dbutil.h"
#include "gpumotif.h"

extern int rm_error;
extern ARGS_T *rm_args;
extern FILE *rm_dbfp;
extern STREL_T *rm_lctx, *rm_rctx;

extern int gm_rm_compile(int, char *[]);
extern int gm_flatten_plan(gm_plan_t *, char *, size_t);
extern int gm_flatten_score(gm_score_t *);
extern void GM_replay_strand(char[], char[], int, int, char[]);
extern int GM_replay_hit_to(const gm_hit_hdr_t *, const gm_hit_el_t *, FILE *);

#define MAX_GPUS 16
#define MAX_WORKERS (2 * MAX_GPUS)

static double now_s(void)
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

static void die_gm(const char *what)
{
	fprintf(stderr, "rnamotif_gpu: %s: %s\n", what, gm_last_error());
	exit(1);
}

static void *xrealloc(void *p, size_t n)
{
	p = realloc(p, n ? n : 1);
	if (p == NULL) {
		fprintf(stderr, "rnamotif_gpu: out of memory\n");
		exit(1);
	}
	return p;
}

/* ------------------------------------------------------------------ run-wide state */

static gm_plan_t plan;
static gm_score_t score;  /* the MAIN score program for the device's pre-screen (gm_ctx_set_score) */
static uint64_t tot_rejected;
static int chk_both_strs, show_progress, stats, prune;
static const char *prog;
static int devices[MAX_GPUS], n_gpus;
static int64_t batch_nt = 128ll << 20;
static int ctx_lead, ctx_trail; /* characters the sink's tail may read before / behind a window */
static double t_read, t_search, t_replay; /* GPUMOTIF_STATS summary */
static double t_setup, t_loop, t_down;
static uint64_t tot_nt, tot_hits;

/* ------------------------------------------------------------------ replay of one batch
 *
 * A batch as the replay sees it: candidates in enumeration order, and for every
 * candidate the name, definition line and length of its record and the characters of
 * the searched strand. */
typedef struct {
	const char *hits;
	size_t n_hits, stride;
	int n_rec;
	const int64_t *rec_off;
	/* host reader: all characters of the batch, names per record */
	char *chars;
	char **sid, **sdef;
	/* device reader: the batch's text, where each record's header starts, one window per candidate */
	const char *text;
	size_t text_len;
	const int64_t *hdr_off;
	const char *wins;
	size_t win_stride;
} BATCH_T;

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

/* sid and sdef of a header line the way FN_fgetseq takes them (src/dbutil.c:62-97).
 * Returns 0 for an entry without a name. */
static int parse_header(const char *p, const char *end, char sid[SID_SIZE], char sdef[SDEF_SIZE])
{
	char *dp;
	unsigned cnt;
	sid[0] = sdef[0] = '\0';
	p++; /* '>' */
	while (p < end && isspace((unsigned char)*p) && *p != '\n')
		p++;
	if (p >= end || *p == '\n')
		return 0;
	for (dp = sid; p < end && !isspace((unsigned char)*p); p++)
		if (dp - sid < SID_SIZE - 1)
			*dp++ = *p;
	*dp = '\0';
	if (p >= end || *p == '\n')
		return 1;
	while (p < end && isspace((unsigned char)*p) && *p != '\n')
		p++;
	if (p >= end || *p == '\n')
		return 1;
	for (dp = sdef, cnt = 0; p < end && *p != '\n'; p++)
		if (++cnt < SDEF_SIZE)
			*dp++ = *p;
	*dp = '\0';
	if (cnt >= SDEF_SIZE)
		fprintf(stderr, "FN_fgetseq: entry: '%s': def len: %d, truncated to %d.\n", sid, cnt, SDEF_SIZE - 1);
	return 1;
}

static void replay_batch(const BATCH_T *b)
{
	static char sid[SID_SIZE], sdef[SDEF_SIZE];
	static char *rcbuf;
	static int rc_cap;
	/* GPUMOTIF_PRUNE: accepted hits of the batch, their text and their records */
	FILE *cap = NULL;
	char *cap_buf = NULL;
	size_t cap_len = 0, n_acc = 0;
	static size_t acc_cap;
	static size_t *acc_end;          /* end of hit k's text in the capture */
	static char *acc_hits;           /* their hit records, contiguous */
	static int32_t *acc_group;       /* rmprune's blocks: locus name up to the first '.' */
	static uint8_t *acc_keep;
	static char (*acc_name)[SID_SIZE];
	size_t h;
	const size_t stride = b->stride;

	if (prune) {
		cap = open_memstream(&cap_buf, &cap_len);
		if (cap == NULL) {
			fprintf(stderr, "rnamotif_gpu: open_memstream failed\n");
			exit(1);
		}
	}
	for (h = 0; h < b->n_hits;) {
		const gm_hit_hdr_t *hdr = (const gm_hit_hdr_t *)(b->hits + h * stride);
		const uint32_t rec = hdr->rec;
		const int comp = hdr->comp;
		const int slen = (int)(b->rec_off[rec + 1] - b->rec_off[rec]);
		char *sb = NULL, saved = 0, *rsid, *rsdef;
		if (b->chars != NULL) {
			/* host reader: the whole strand is at hand */
			rsid = b->sid[rec];
			rsdef = b->sdef[rec];
			sb = b->chars + b->rec_off[rec];
			saved = sb[slen];
			sb[slen] = '\0';
			if (comp) {
				if (slen + 1 > rc_cap) {
					rc_cap = slen + 1;
					rcbuf = xrealloc(rcbuf, rc_cap);
				}
				revcomp_into(sb, slen, rcbuf);
				GM_replay_strand(rsid, rsdef, 1, slen, rcbuf);
			} else
				GM_replay_strand(rsid, rsdef, 0, slen, sb);
		} else {
			parse_header(b->text + b->hdr_off[rec], b->text + b->hdr_off[rec + 1], sid, sdef);
			rsid = sid;
			rsdef = sdef;
		}
		for (; h < b->n_hits; h++) {
			hdr = (const gm_hit_hdr_t *)(b->hits + h * stride);
			if (hdr->rec != rec || hdr->comp != comp)
				break;
			if (b->chars == NULL) {
				/* device reader: window h holds the strand from szero - lead on, so the
				 * strand buffer "starts" szero - lead characters before it */
				char *w = (char *)b->wins + h * b->win_stride;
				GM_replay_strand(rsid, rsdef, comp, slen, w - ((long)hdr->szero - ctx_lead));
			}
			if (!prune) {
				GM_replay_hit_to(hdr, (const gm_hit_el_t *)(hdr + 1), stdout);
				continue;
			}
			if (GM_replay_hit_to(hdr, (const gm_hit_el_t *)(hdr + 1), cap)) {
				fflush(cap);
				if (n_acc == acc_cap) {
					acc_cap = acc_cap ? 2 * acc_cap : 1024;
					acc_end = xrealloc(acc_end, acc_cap * sizeof *acc_end);
					acc_hits = xrealloc(acc_hits, acc_cap * stride);
					acc_group = xrealloc(acc_group, acc_cap * sizeof *acc_group);
					acc_keep = xrealloc(acc_keep, acc_cap);
					acc_name = xrealloc(acc_name, acc_cap * sizeof *acc_name);
				}
				acc_end[n_acc] = cap_len;
				memcpy(acc_hits + n_acc * stride, hdr, stride);
				snprintf(acc_name[n_acc], SID_SIZE, "%s", rsid);
				n_acc++;
			}
		}
		if (sb != NULL)
			sb[slen] = saved;
	}
	if (prune) {
		size_t k, from = 0;
		int32_t gid = 0;
		fclose(cap);
		/* rmprune's blocks are runs of hits whose locus names agree up to the
		 * first '.' (getname, src/rmprune.c:312-330): number them */
		for (k = 0; k < n_acc; k++) {
			if (k > 0) {
				const char *a = acc_name[k], *bb = acc_name[k - 1];
				size_t la = strcspn(a, "."), lb = strcspn(bb, ".");
				if (la != lb || strncmp(a, bb, la))
					gid++;
			}
			acc_group[k] = gid;
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
	}
	tot_hits += b->n_hits;
}

/* ------------------------------------------------------------------ host reader
 *
 * The reference's own fgetseq (any format), one batch at a time: read, upload the
 * characters, search, replay.  Starts at the current position of rm_dbfp. */
static void run_host_reader(gm_ctx *ctx, int ecnt)
{
	static char sid[SID_SIZE], sdef[SDEF_SIZE];
	int (*fgetseq)(FILE *, char *, int, char *, int, char *);
	int64_t buf_cap, used;
	char *buf;
	char **sids = NULL, **sdefs = NULL;
	int64_t *offs = NULL;
	int n_recs, cap_recs = 0, eof = 0, r;

	if (rm_args->a_dbfmt == NULL || !strcmp(rm_args->a_dbfmt, DT_FASTN))
		fgetseq = FN_fgetseq;
	else if (!strcmp(rm_args->a_dbfmt, DT_PIR))
		fgetseq = PIR_fgetseq;
	else
		fgetseq = GB_fgetseq;
	buf_cap = batch_nt + rm_args->a_maxslen + 16;
	buf = malloc((size_t)buf_cap);
	if (buf == NULL) {
		rm_error = TRUE;
		LOG_ERROR("can't allocate sbuf (s_sbuf=%lld)", (long long)buf_cap);
		exit(1);
	}
	while (!eof) {
		BATCH_T b;
		const void *hp;
		double t0 = now_s(), t1, t2;

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
				fprintf(stderr, "%s: %7d: %s\n", prog, ecnt, sid);
			if (n_recs == cap_recs) {
				cap_recs = cap_recs ? 2 * cap_recs : 1024;
				sids = xrealloc(sids, cap_recs * sizeof *sids);
				sdefs = xrealloc(sdefs, cap_recs * sizeof *sdefs);
				offs = xrealloc(offs, (cap_recs + 1) * sizeof *offs);
			}
			sids[n_recs] = strdup(sid);
			sdefs[n_recs] = strdup(sdef);
			offs[n_recs] = used;
			n_recs++;
			used += slen; /* the NUL fgetseq wrote is overwritten by the next record;
				       * the replay restores it per record */
		}
		if (n_recs == 0)
			break;
		offs[n_recs] = used;
		t1 = now_s();
		if (gm_db_upload_chars(ctx, buf, offs, n_recs))
			die_gm("gm_db_upload_chars");
		if (gm_scan(ctx, 0, used, chk_both_strs ? 2 : 1))
			die_gm("gm_scan");
		memset(&b, 0, sizeof b);
		if (gm_hits(ctx, &hp, &b.n_hits, &b.stride))
			die_gm("gm_hits");
		{
			gm_scan_stats_t st;
			gm_stats(ctx, &st);
			tot_rejected += st.n_score_rejected;
		}
		t2 = now_s();
		b.hits = hp;
		b.n_rec = n_recs;
		b.rec_off = offs;
		b.chars = buf;
		b.sid = sids;
		b.sdef = sdefs;
		replay_batch(&b);
		t_read += t1 - t0;
		t_search += t2 - t1;
		t_replay += now_s() - t2;
		tot_nt += (uint64_t)used;
		if (stats)
			fprintf(stderr, "rnamotif_gpu: host reader: batch of %d records, %lld nt: read %.1f ms, search %.1f ms, "
				"replay %.1f ms, %zu candidates\n", n_recs, (long long)used, 1e3 * (t1 - t0), 1e3 * (t2 - t1),
				1e3 * (now_s() - t2), b.n_hits);
		for (r = 0; r < n_recs; r++) {
			free(sids[r]);
			free(sdefs[r]);
		}
	}
	free(buf);
}

/* ------------------------------------------------------------------ the pipeline */

enum { SL_FREE = 0, SL_BUSY, SL_READY, SL_BAIL };

typedef struct {
	gm_ctx *ctx;
	int device;
	char *text;              /* pinned */
	size_t text_cap, text_len;
	long seq;                /* number of the batch in the slot */
	int state;
	int file;                /* where the batch came from: file index, offset of its first byte */
	off_t file_off;
	BATCH_T b;
	double ms_read, ms_search;
	uint64_t n_rejected;     /* candidates the score pre-screen dropped on the device */
	pthread_t thread;
} SLOT_T;

static SLOT_T slots[MAX_WORKERS];
static int n_slots;
static pthread_mutex_t mu = PTHREAD_MUTEX_INITIALIZER;
static pthread_cond_t cv = PTHREAD_COND_INITIALIZER;
/* the reader (under mu): which file, where in it, the next batch number */
static int rd_file = -1, rd_fd = -1, rd_done, rd_stop;
static off_t rd_off, rd_size;
static long rd_seq;
static int rd_turn_free = 1; /* one worker reads at a time, in batch order */

/* pread of a large range with several threads (one thread moves ~6 GB/s out of the
 * page cache; the device searches ten times that) */
#define READ_THREADS 8
typedef struct {
	int fd;
	char *dst;
	size_t n;
	off_t off;
	size_t got;
} READ_PART_T;

static void *read_part(void *arg)
{
	READ_PART_T *p = arg;
	while (p->got < p->n) {
		ssize_t k = pread(p->fd, p->dst + p->got, p->n - p->got, p->off + (off_t)p->got);
		if (k <= 0)
			break;
		p->got += (size_t)k;
	}
	return NULL;
}

/* returns the number of bytes read from the front of the range without a gap */
static size_t parallel_pread(int fd, char *dst, size_t n, off_t off)
{
	READ_PART_T part[READ_THREADS];
	pthread_t th[READ_THREADS];
	int k, nt = n < ((size_t)8 << 20) ? 1 : READ_THREADS;
	size_t per = (n / (size_t)nt + 4095) & ~(size_t)4095, done = 0;
	for (k = 0; k < nt; k++) {
		const size_t lo = per * (size_t)k < n ? per * (size_t)k : n;
		const size_t hi = k == nt - 1 || lo + per > n ? n : lo + per;
		part[k].fd = fd;
		part[k].dst = dst + lo;
		part[k].n = hi - lo;
		part[k].off = off + (off_t)lo;
		part[k].got = 0;
		if (k == 0 || pthread_create(&th[k], NULL, read_part, &part[k]))
			th[k] = 0;
	}
	read_part(&part[0]);
	for (k = 1; k < nt; k++) {
		if (th[k])
			pthread_join(th[k], NULL);
		else
			read_part(&part[k]);
	}
	for (k = 0; k < nt; k++) {
		done += part[k].got;
		if (part[k].got < part[k].n)
			break;
	}
	return done;
}

/* Read the next batch -- whole records: from a '>' that starts a line to the last
 * "\n>" inside the window, or the end of the file -- into the slot's pinned buffer.
 * Returns 0 at the end of the input.  The first byte of a file must be '>'; anything
 * else is left to the host reader (SL_BAIL). */
static int read_batch(SLOT_T *s)
{
	for (;;) {
		size_t want, got = 0;
		if (rd_fd < 0) {
			struct stat st;
			if (rd_file + 1 >= rm_args->a_n_dbfname)
				return 0;
			rd_file++;
			rd_fd = open(rm_args->a_dbfname[rd_file], O_RDONLY);
			if (rd_fd < 0 || fstat(rd_fd, &st) || !S_ISREG(st.st_mode)) {
				/* unreadable or not a plain file: the host reader reports / handles it */
				if (rd_fd >= 0)
					close(rd_fd);
				rd_fd = -1;
				s->file = rd_file;
				s->file_off = 0;
				s->text_len = 0;
				return -1;
			}
			rd_off = 0;
			rd_size = st.st_size;
		}
		if (rd_off >= rd_size) {
			close(rd_fd);
			rd_fd = -1;
			continue;
		}
		s->file = rd_file;
		s->file_off = rd_off;
		want = (size_t)batch_nt;
		for (;;) {
			size_t cut;
			if (want > (size_t)(rd_size - rd_off))
				want = (size_t)(rd_size - rd_off);
			if (want + 16 > s->text_cap) {
				void *p = NULL;
				if (gm_host_alloc(&p, want + want / 8 + 4096))
					die_gm("gm_host_alloc");
				if (got)
					memcpy(p, s->text, got);
				gm_host_free(s->text);
				s->text = p;
				s->text_cap = want + want / 8 + 4096;
			}
			if (got < want) {
				const size_t k = parallel_pread(rd_fd, s->text + got, want - got, rd_off + (off_t)got);
				got += k;
				if (got < want) {
					/* the file shrank or cannot be read: take what there is */
					rd_size = rd_off + (off_t)got;
					want = got;
				}
			}
			if (got == 0)
				break;
			if (s->text[0] != '>')
				return -1; /* "fastn file does not begin with '>'" (src/dbutil.c:56-60) */
			if (rd_off + (off_t)got >= rd_size) {
				cut = got; /* to the end of the file */
			} else {
				/* the last '>' at the start of a line */
				cut = got;
				while (cut > 1) {
					const char *q = memrchr(s->text + 1, '>', cut - 1);
					if (q == NULL) {
						cut = 0;
						break;
					}
					cut = (size_t)(q - s->text);
					if (s->text[cut - 1] == '\n')
						break;
				}
				if (cut <= 1) {
					/* one record fills the window: widen it */
					want = got * 2;
					continue;
				}
			}
			s->text_len = cut;
			rd_off += (off_t)cut;
			return 1;
		}
		close(rd_fd);
		rd_fd = -1;
	}
}

static void *worker(void *arg)
{
	SLOT_T *s = arg;
	for (;;) {
		double t0, t1;
		int rc, r, bail = 0;
		pthread_mutex_lock(&mu);
		while (s->state != SL_FREE && !rd_stop)
			pthread_cond_wait(&cv, &mu);
		while (!rd_turn_free && !rd_stop)
			pthread_cond_wait(&cv, &mu);
		if (rd_stop || rd_done) {
			pthread_mutex_unlock(&mu);
			return NULL;
		}
		rd_turn_free = 0;
		s->seq = rd_seq++;
		s->state = SL_BUSY;
		pthread_mutex_unlock(&mu);

		t0 = now_s();
		rc = read_batch(s); /* rd_* belong to the worker whose turn it is */
		t1 = now_s();

		pthread_mutex_lock(&mu);
		if (rc == 0) {
			rd_done = 1;
			rd_seq = s->seq; /* this number was not used */
			s->state = SL_FREE;
		} else if (rc < 0) {
			rd_done = 1;
			rd_seq = s->seq + 1;
			s->state = SL_BAIL;
		}
		rd_turn_free = 1;
		pthread_cond_broadcast(&cv);
		pthread_mutex_unlock(&mu);
		if (rc <= 0)
			return NULL;

		/* ---- search ---- */
		memset(&s->b, 0, sizeof s->b);
		if (gm_db_upload_fastn(s->ctx, s->text, s->text_len))
			die_gm("gm_db_upload_fastn");
		if (gm_db_records(s->ctx, &s->b.rec_off, &s->b.hdr_off, &s->b.n_rec))
			die_gm("gm_db_records");
		/* what FN_fgetseq would not read the same way goes to the host reader */
		for (r = 0; r < s->b.n_rec && !bail; r++) {
			const char *p = s->text + s->b.hdr_off[r] + 1, *e = s->text + s->b.hdr_off[r + 1];
			while (p < e && isspace((unsigned char)*p) && *p != '\n')
				p++;
			if (p >= e || *p == '\n')
				bail = 1; /* an entry without a name ends the file (:62-66) */
			if (s->b.rec_off[r + 1] - s->b.rec_off[r] > (int64_t)rm_args->a_maxslen - 1)
				bail = 1; /* longer than -maxslen: truncated there (:104-125) */
		}
		if (!bail) {
			const void *hp;
			if (gm_scan(s->ctx, 0, gm_db_total_nt(s->ctx), chk_both_strs ? 2 : 1))
				die_gm("gm_scan");
			if (gm_hits(s->ctx, &hp, &s->b.n_hits, &s->b.stride))
				die_gm("gm_hits");
			s->b.hits = hp;
			{
				gm_scan_stats_t st;
				gm_stats(s->ctx, &st);
				s->n_rejected = st.n_score_rejected;
			}
			if (s->b.n_hits > 0 && gm_hit_windows(s->ctx, ctx_lead, ctx_trail, &s->b.wins, &s->b.win_stride))
				die_gm("gm_hit_windows");
			s->b.text = s->text;
			s->b.text_len = s->text_len;
		}
		s->ms_read = 1e3 * (t1 - t0);
		s->ms_search = 1e3 * (now_s() - t1);

		pthread_mutex_lock(&mu);
		s->state = bail ? SL_BAIL : SL_READY;
		if (bail)
			rd_stop = 1;
		pthread_cond_broadcast(&cv);
		pthread_mutex_unlock(&mu);
		if (bail)
			return NULL;
	}
}

/* Returns the number of records replayed; *bail_file / *bail_off tell where the host
 * reader has to take over (file index -1: nowhere, the input is done). */
static int run_pipeline(int *bail_file, off_t *bail_off)
{
	long seq;
	int i, ecnt = 0;
	static char sid[SID_SIZE], sdef[SDEF_SIZE];

	*bail_file = -1;
	*bail_off = 0;
	for (i = 0; i < n_slots; i++)
		if (pthread_create(&slots[i].thread, NULL, worker, &slots[i])) {
			fprintf(stderr, "rnamotif_gpu: pthread_create failed\n");
			exit(1);
		}
	for (seq = 0;; seq++) {
		SLOT_T *s = NULL;
		double t0, w0 = now_s();
		pthread_mutex_lock(&mu);
		for (;;) {
			for (i = 0; i < n_slots; i++)
				if (slots[i].seq == seq && (slots[i].state == SL_READY || slots[i].state == SL_BAIL))
					s = &slots[i];
			if (s != NULL || (rd_done && seq >= rd_seq))
				break;
			pthread_cond_wait(&cv, &mu);
		}
		pthread_mutex_unlock(&mu);
		if (s == NULL)
			break; /* all batches replayed */
		if (s->state == SL_BAIL) {
			*bail_file = s->file;
			*bail_off = s->file_off;
			pthread_mutex_lock(&mu);
			rd_stop = 1;
			pthread_cond_broadcast(&cv);
			pthread_mutex_unlock(&mu);
			break;
		}
		t0 = now_s();
		if (show_progress) {
			int r;
			for (r = 0; r < s->b.n_rec; r++)
				if (++ecnt % show_progress == 0) {
					parse_header(s->text + s->b.hdr_off[r], s->text + s->b.hdr_off[r + 1], sid, sdef);
					fprintf(stderr, "%s: %7d: %s\n", prog, ecnt, sid);
				}
		} else
			ecnt += s->b.n_rec;
		replay_batch(&s->b);
		t_search += t0 - w0; /* time the replay waited for a batch */
		t_replay += now_s() - t0;
		tot_nt += (uint64_t)s->b.rec_off[s->b.n_rec];
		tot_rejected += s->n_rejected;
		if (stats)
			fprintf(stderr, "rnamotif_gpu: batch %ld on gpu %d: %d records, %lld nt: read %.1f ms, search %.1f ms, "
				"replay %.1f ms (waited %.1f ms), %zu candidates\n", seq, s->device, s->b.n_rec,
				(long long)s->b.rec_off[s->b.n_rec], s->ms_read, s->ms_search, 1e3 * (now_s() - t0), 1e3 * (t0 - w0),
				s->b.n_hits);
		pthread_mutex_lock(&mu);
		s->state = SL_FREE;
		s->seq = -1;
		pthread_cond_broadcast(&cv);
		pthread_mutex_unlock(&mu);
	}
	for (i = 0; i < n_slots; i++)
		pthread_join(slots[i].thread, NULL);
	if (rd_fd >= 0)
		close(rd_fd);
	return ecnt;
}

int main(int argc, char *argv[])
{
	char err[512];
	IDENT_T *ip;
	const char *ev;
	int g, i, host_only = 0, ecnt = 0;
	double t_start = now_s();
	int fmt = -1, saved_stdout = -1; /* GPUMOTIF_FMT */
	FILE *fmt_tmp = NULL;

	prog = argv[0];
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
	if (n_gpus == 0)
		devices[n_gpus++] = (ev = getenv("GPUMOTIF_DEVICE")) != NULL ? atoi(ev) : 0;
	if ((ev = getenv("GPUMOTIF_BATCH_NT")) != NULL && atoll(ev) > 0)
		batch_nt = atoll(ev);
	if ((ev = getenv("GPUMOTIF_STATS")) != NULL)
		stats = atoi(ev);
	if ((ev = getenv("GPUMOTIF_PRUNE")) != NULL)
		prune = atoi(ev);
	if ((ev = getenv("GPUMOTIF_READER")) != NULL && !strcmp(ev, "host"))
		host_only = 1;
	if ((ev = getenv("GPUMOTIF_FMT")) != NULL && *ev)
		fmt = !strcmp(ev, "la") ? 2 : !strcmp(ev, "l") ? 1 : 0;
	if (fmt >= 0) {
		/* everything written to stdout from here on (print_match, the score program's own
		 * output) goes to a temporary file first */
		fflush(stdout);
		saved_stdout = dup(1);
		fmt_tmp = tmpfile();
		if (saved_stdout < 0 || fmt_tmp == NULL || dup2(fileno(fmt_tmp), 1) < 0) {
			fprintf(stderr, "rnamotif_gpu: GPUMOTIF_FMT: can't redirect stdout\n");
			exit(1);
		}
	}

	ip = RM_find_id("chk_both_strs");
	chk_both_strs = ip == NULL ? 1 : ip->i_val.v_value.v_ival;
	ip = RM_find_id("show_progress");
	show_progress = ip == NULL ? 0 : ip->i_val.v_value.v_ival;

	/* (the reference's main() has checked the format name and opened the first file: rm_dbfp) */
	if (rm_args->a_dbfmt != NULL && strcmp(rm_args->a_dbfmt, DT_FASTN))
		host_only = 1;
	if (rm_args->a_n_dbfname == 0)
		host_only = 1; /* standard input */
	/* the sink's tail reads the context around a match when there is one (set_context,
	 * src/find_motif.c:1720-1756; print_match :1853-1862) */
	ctx_lead = plan.lctx.present ? plan.lctx.maxlen : 0;
	ctx_trail = plan.rctx.present ? 2 * plan.rctx.maxlen : 0;

	t_setup = now_s() - t_start; /* front end + plan */
	if (RM_fm_init())
		exit(1);
	RM_setprog(P_BEGIN);
	RM_score(0, 0, NULL, NULL);
	RM_setprog(P_MAIN);
	/* the score program as the BEGIN section left it: candidates it rejects outright are
	 * dropped on the device and never replayed (GPUMOTIF_NO_SCORE=1: replay everything) */
	gm_flatten_score(&score);
	if (prune && score.has_hold) {
		/* held hits are printed by score.c itself, later and past the capture: their order
		 * would change and they would escape the pruning */
		fprintf(stderr, "rnamotif_gpu: GPUMOTIF_PRUNE is not available for score programs that HOLD / RELEASE hits; "
			"pipe the output through rmprune instead\n");
		exit(1);
	}
	if (getenv("GPUMOTIF_NO_SCORE") != NULL)
		score.present = 0;
	if (stats)
		fprintf(stderr, "rnamotif_gpu: score pre-screen on the device: %s%s%s%s\n", score.present ? "on" : "off",
			score.present || !score.why[0] ? "" : " (", score.present ? "" : score.why,
			score.present || !score.why[0] ? "" : ")");

	if (!host_only) {
		int bail_file;
		off_t bail_off;
		n_slots = 2 * n_gpus;
		for (i = 0; i < n_slots; i++) {
			slots[i].device = devices[i % n_gpus];
			slots[i].seq = -1;
			if (gm_ctx_create(&slots[i].ctx, &plan, slots[i].device))
				die_gm("gm_ctx_create");
			if (score.present && gm_ctx_set_score(slots[i].ctx, &score))
				die_gm("gm_ctx_set_score");
		}
		{
			const double t0 = now_s();
			t_setup = t0 - t_start; /* ... + contexts */
			ecnt = run_pipeline(&bail_file, &bail_off);
			t_loop = now_s() - t0;
		}
		if (bail_file >= 0) {
			/* hand over to the host reader at the batch the pipeline stopped at; a file
			 * that cannot be opened ends the run there like DB_fnext does (src/dbutil.c:25-38) */
			if (rm_dbfp != NULL && rm_dbfp != stdin)
				fclose(rm_dbfp);
			rm_args->a_c_dbfname = bail_file;
			rm_dbfp = fopen(rm_args->a_dbfname[bail_file], "r");
			if (rm_dbfp == NULL)
				fprintf(stderr, "DB_fnext: can't read seq file '%s'.\n", rm_args->a_dbfname[bail_file]);
			else {
				if (bail_off > 0)
					fseeko(rm_dbfp, bail_off, SEEK_SET);
				run_host_reader(slots[0].ctx, ecnt);
			}
		}
		{
			const double t0 = now_s();
			for (i = 0; i < n_slots; i++) {
				gm_ctx_destroy(slots[i].ctx);
				gm_host_free(slots[i].text);
			}
			t_down = now_s() - t0;
		}
	} else {
		gm_ctx *ctx;
		if (gm_ctx_create(&ctx, &plan, devices[0]))
			die_gm("gm_ctx_create");
		if (score.present && gm_ctx_set_score(ctx, &score))
			die_gm("gm_ctx_set_score");
		run_host_reader(ctx, 0); /* from the start of the first file (or standard input) */
		gm_ctx_destroy(ctx);
	}

	RM_setprog(P_END);
	RM_score(0, 0, NULL, NULL);
	if (fmt >= 0) {
		char *text;
		off_t n;
		fflush(stdout);
		n = lseek(1, 0, SEEK_END);
		text = n > 0 ? malloc((size_t)n) : NULL;
		if (n > 0 && (text == NULL || pread(1, text, (size_t)n, 0) != (ssize_t)n)) {
			fprintf(stderr, "rnamotif_gpu: GPUMOTIF_FMT: can't read the collected output back\n");
			exit(1);
		}
		dup2(saved_stdout, 1);
		close(saved_stdout);
		if (gm_rmfmt(text, n > 0 ? (size_t)n : 0, fmt, stdout))
			die_gm("gm_rmfmt");
		free(text);
	}
	if (stats) {
		const double wall = now_s() - t_start;
		const double snt = (double)tot_nt * (chk_both_strs ? 2 : 1);
		fprintf(stderr, "rnamotif_gpu: score pre-screen dropped %llu candidates on the device\n", (unsigned long long)tot_rejected);
		fprintf(stderr, "rnamotif_gpu: %llu nt, %llu candidates replayed, %d gpu(s), wall %.3f s "
			"(%.2f G strand-nt/s): set-up %.3f s, pipeline %.3f s (%.2f G strand-nt/s; replay %.3f s, replay waiting "
			"for batches %.3f s), tear-down %.3f s, host reader %.3f s\n",
			(unsigned long long)tot_nt, (unsigned long long)tot_hits, n_gpus, wall, snt / wall / 1e9, t_setup, t_loop,
			t_loop > 0 ? snt / t_loop / 1e9 : 0.0, t_replay, t_search, t_down, t_read);
	}
	(void)g;
	exit(0);
}
