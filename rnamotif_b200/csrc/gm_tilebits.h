// gm_tilebits.h -- eight nucleotides at a time: from a word of eight packed 4-bit
// IUPAC codes to what a staged tile holds for them -- the expanded byte of each
// (low nibble the code, high nibble the reference's base code rm_b2bc: a c g t = 0..3,
// anything else 4), the bytes of the reverse complement (mk_rcmp,
// src/rnamot.c:200-208: a<->t, c<->g, anything else n), and the eight bits of each
// of the four base bitsets.  SIMD within a register, no branches, no table: the tile
// loader of gm_machine.cuh runs it once per lane per 256 nucleotides.  Plain integer
// code, so the host build (tests/test_tilebits.py) checks it against the per-nucleotide
// definitions on every code word.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define GM_HD __host__ __device__ __forceinline__
#else
#define GM_HD static inline
#endif

namespace gm {

// nibbles 0..3 of the low half of x spread to the low nibbles of four bytes
GM_HD uint32_t spread4(uint32_t h16)
{
	uint32_t t = (h16 | (h16 << 8)) & 0x00ff00ffu;
	return (t | (t << 4)) & 0x0f0f0f0fu;
}

// bits 0, 4, 8, 12 of y (nothing else set) to bits 0..3: the sixteen partial products of
// y * 0x249 fall on sixteen different bits, so nothing carries
GM_HD uint32_t gather4(uint32_t y)
{
	return ((y * 0x249u) >> 9) & 0xfu;
}

// one half (four nucleotides): t = their codes, one per byte; hb = 0x01 in the bytes whose
// code is a single base.  fwd = expanded bytes in position order, rc = complement bytes in
// REVERSED position order (byte 0 = the last nucleotide)
GM_HD void expand4(uint32_t t, uint32_t hb, uint32_t &fwd, uint32_t &rc)
{
	const uint32_t m7 = hb * 7u, mff = hb * 255u;
	// log2 of a one-hot nibble: (c >> 1) - (c >> 3); 0 where the code is not a single base
	const uint32_t lg = (((t >> 1) & 0x07070707u) - ((t >> 3) & 0x01010101u)) & m7;
	const uint32_t b = lg | ((hb ^ 0x01010101u) << 2); // base code, 4 for anything else
	fwd = t | (b << 4);
	// complement of a single base: the code with its four bits reversed, base code 3 - b
	const uint32_t rv = ((t & 0x01010101u) << 3) | ((t & 0x02020202u) << 1) | ((t >> 1) & 0x02020202u) | ((t >> 3) & 0x01010101u);
	const uint32_t nb = 0x03030303u - lg;
	const uint32_t c = ((rv | (nb << 4)) & mff) | (0x4f4f4f4fu & ~mff);
	rc = (c >> 24) | ((c >> 8) & 0xff00u) | ((c << 8) & 0xff0000u) | (c << 24);
}

// x: eight codes, nucleotide j in nibble j.
//   f0 f1   expanded bytes of nucleotides 0..3 and 4..7
//   r0 r1   complement bytes of nucleotides 7 6 5 4 and 3 2 1 0 (ascending addresses of the
//           reverse-complement strand)
//   bits    byte x = the eight bits of base x's bitset (bit j: nucleotide j is base x)
GM_HD void expand8(uint32_t x, uint32_t &f0, uint32_t &f1, uint32_t &r0, uint32_t &r1, uint32_t &bits)
{
	const uint32_t M = 0x11111111u;
	const uint32_t b0 = x & M, b1 = (x >> 1) & M, b2 = (x >> 2) & M, b3 = (x >> 3) & M;
	const uint32_t s = b0 + b1 + b2 + b3;        // bits set per nibble, 0..4
	const uint32_t h = s & ~(s >> 1) & M;         // exactly one
	const uint32_t a = b0 & h, c = b1 & h, g = b2 & h, t = b3 & h;
	bits = gather4(a & 0xffffu) | (gather4(a >> 16) << 4) | (gather4(c & 0xffffu) << 8) | (gather4(c >> 16) << 12) |
		(gather4(g & 0xffffu) << 16) | (gather4(g >> 16) << 20) | (gather4(t & 0xffffu) << 24) | (gather4(t >> 16) << 28);
	uint32_t rlo, rhi;
	expand4(spread4(x & 0xffffu), spread4(h & 0xffffu), f0, rlo);
	expand4(spread4(x >> 16), spread4(h >> 16), f1, rhi);
	r0 = rhi;
	r1 = rlo;
}

} // namespace gm
