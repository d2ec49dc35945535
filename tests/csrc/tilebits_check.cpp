// Host check of gm_tilebits.h against the per-nucleotide definitions of the tile loader
// (expand_code / complement_byte, rnamotif_b200/csrc/gm_machine.cuh) on every 16-bit code
// half in both halves of the word and on random words.  Prints "ok" or the first mismatch.
#include <cstdio>
#include <cstdlib>
#include "gm_tilebits.h"

static uint8_t expand_code(unsigned c)
{
	unsigned b = c == 1 ? 0 : c == 2 ? 1 : c == 4 ? 2 : c == 8 ? 3 : 4;
	return (uint8_t)(c | (b << 4));
}
static uint8_t complement_byte(uint8_t v)
{
	unsigned b = v >> 4;
	if (b > 3)
		return (uint8_t)(15 | (4 << 4));
	unsigned nb = 3 - b;
	return (uint8_t)((1u << nb) | (nb << 4));
}

static int check(uint32_t x)
{
	uint32_t f[2], r[2], bits;
	gm::expand8(x, f[0], f[1], r[0], r[1], bits);
	const uint8_t *fb = (const uint8_t *)f, *rb = (const uint8_t *)r;
	uint32_t want = 0;
	for (int j = 0; j < 8; j++) {
		const uint8_t v = expand_code((x >> (4 * j)) & 15);
		if (fb[j] != v || rb[7 - j] != complement_byte(v)) {
			printf("mismatch: word %08x nucleotide %d: fwd %02x (want %02x) rc %02x (want %02x)\n", x, j, fb[j], v, rb[7 - j],
			       complement_byte(v));
			return 1;
		}
		if ((v >> 4) < 4)
			want |= 1u << (8 * (v >> 4) + j);
	}
	if (bits != want) {
		printf("mismatch: word %08x bitset bytes %08x (want %08x)\n", x, bits, want);
		return 1;
	}
	return 0;
}

int main()
{
	for (uint32_t h = 0; h < 65536; h++)
		if (check(h) || check(h << 16) || check(h * 0x10001u) || check(h | (~h << 16)))
			return 1;
	srand(7);
	for (int i = 0; i < 2000000; i++)
		if (check((uint32_t)rand() ^ ((uint32_t)rand() << 11) ^ ((uint32_t)rand() << 22)))
			return 1;
	puts("ok");
	return 0;
}
