/*
 * malloc_ff.c -- TEST INFRASTRUCTURE (LD_PRELOAD shim for the reference binary).
 *
 * The reference reads fm_window[] cells it has never written: the array comes
 * from malloc() (src/find_motif.c:129) and only the two cells around the
 * window are initialised per start (:191-192), so chk_wchlx & co
 * (:1460-1703) see whatever the heap held for positions outside the current
 * match that no earlier candidate has marked and unmarked.  With glibc that
 * is usually 0 (= element 0), i.e. the verdict of -strict_helices depends on
 * the allocator and on every record scanned before.  Preloading this shim
 * hands the reference memory filled with 0xff, so a never-written cell reads
 * UNDEF (-1) -- the value unmark_*() leaves behind -- and the candidate
 * stream becomes a function of the input alone.  The goldens under
 * tests/golden/ are generated this way (tests/golden/make_golden.py).
 */
#include <stddef.h>
#include <string.h>

extern void *__libc_malloc(size_t);

void *malloc(size_t n)
{
	void *p = __libc_malloc(n);
	if (p != NULL)
		memset(p, 0xff, n);
	return p;
}
