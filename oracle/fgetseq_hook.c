/*
 * fgetseq_hook.c -- TEST INFRASTRUCTURE.  Runs the REFERENCE's own FN_fgetseq
 * (src/dbutil.c:42-128, compiled from where it lies into oracle/_ref/) over a
 * FASTA text held in memory and returns what it read, so that the device
 * reader (gm_db_upload_fastn) can be compared against it byte for byte.
 *
 *   gmo_ref_fastn(text, n, maxslen, seq, seq_cap, rec_off, ids, ids_cap, max_rec)
 *     -> number of records, or -1 when a buffer is too small.
 *     seq      the records' characters, concatenated (lower case, u -> t)
 *     rec_off  n_rec + 1 offsets into seq
 *     ids      "id\tdef\n" per record
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "rmdefs.h"
#include "dbutil.h"

#ifndef SID_SIZE
#define SID_SIZE 100
#endif
#ifndef SDEF_SIZE
#define SDEF_SIZE 20000
#endif

int gmo_ref_fastn(const char *text, long n, int maxslen, char *seq, long seq_cap, long *rec_off,
	char *ids, long ids_cap, int max_rec)
{
	static char sid[SID_SIZE * 64], sdef[SDEF_SIZE];
	FILE *fp;
	char *sbuf;
	long used = 0, iused = 0;
	int n_rec = 0, slen;

	if (n == 0) {
		rec_off[0] = 0;
		return 0;
	}
	fp = fmemopen((void *)text, (size_t)n, "r");
	sbuf = malloc((size_t)maxslen + 1);
	if (fp == NULL || sbuf == NULL)
		return -1;
	rec_off[0] = 0;
	while ((slen = FN_fgetseq(fp, sid, SDEF_SIZE, sdef, maxslen, sbuf)) != EOF) {
		long need = (long)strlen(sid) + (long)strlen(sdef) + 2;
		if (n_rec >= max_rec || used + slen > seq_cap || iused + need + 1 > ids_cap) {
			n_rec = -1;
			break;
		}
		memcpy(seq + used, sbuf, (size_t)slen);
		used += slen;
		rec_off[++n_rec] = used;
		iused += sprintf(ids + iused, "%s\t%s\n", sid, sdef);
	}
	fclose(fp);
	free(sbuf);
	return n_rec;
}
