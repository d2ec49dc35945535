// gm_hostpack.cpp -- characters -> 4-bit IUPAC codes on the HOST cores, for
// gm_db_upload_chars_hostpack (include/gpumotif.h): a caller whose sequence sits
// in host memory as one character per nucleotide (what FN_fgetseq leaves,
// src/dbutil.c:105-111) then sends half a byte per nucleotide over PCIe instead
// of one.  Same code table as the device's gm_pack_kernel (gm_kernel.cuh,
// code_of_char): case folded, u = t, any other letter 0; nucleotide g goes to
// byte g >> 1, nibble g & 1.
//
// Plain host C++ (no CUDA): AVX2 where the CPU has it (two pshufb look-ups per 32
// characters, pmaddubsw to join the nibbles), a table loop otherwise; a small
// persistent thread team cuts a chunk into 64-nucleotide-aligned pieces.
#include "gm_hostpack.h"

#include <cstring>
#include <cstdlib>
#include <algorithm>
#if defined(__x86_64__)
#include <immintrin.h>
#endif
#include <sched.h>

namespace gm {

namespace {

struct CodeTable {
	uint8_t t[256];
	CodeTable()
	{
		memset(t, 0, sizeof t);
		static const struct { char ch; uint8_t code; } map[] = {
			{'a', 1}, {'c', 2}, {'g', 4}, {'t', 8}, {'u', 8}, {'r', 5}, {'y', 10}, {'m', 3}, {'k', 12},
			{'s', 6}, {'w', 9}, {'h', 11}, {'b', 14}, {'v', 7}, {'d', 13}, {'n', 15}};
		for (const auto &m : map) {
			t[(unsigned char)m.ch] = m.code;
			t[(unsigned char)(m.ch - 32)] = m.code; // upper case
		}
	}
};
const CodeTable g_tab;

// n nucleotides from s into (n + 1) / 2 bytes at out
void pack_scalar(const uint8_t *s, int64_t n, uint8_t *out)
{
	int64_t i = 0;
	for (; i + 1 < n; i += 2)
		out[i >> 1] = (uint8_t)(g_tab.t[s[i]] | (g_tab.t[s[i + 1]] << 4));
	if (i < n)
		out[i >> 1] = g_tab.t[s[i]];
}

#if defined(__x86_64__)
__attribute__((target("avx2"))) inline __m256i codes_avx2(__m256i x, __m256i lut0, __m256i lut1)
{
	const __m256i idx = _mm256_sub_epi8(_mm256_or_si256(x, _mm256_set1_epi8(0x20)), _mm256_set1_epi8(0x61));
	// a letter <=> idx in 0..25 (unsigned)
	const __m256i valid = _mm256_cmpeq_epi8(_mm256_min_epu8(idx, _mm256_set1_epi8(25)), idx);
	const __m256i t0 = _mm256_shuffle_epi8(lut0, idx); // letters a..p by the low four bits
	const __m256i t1 = _mm256_shuffle_epi8(lut1, idx); // letters q..z
	const __m256i hi = _mm256_cmpgt_epi8(idx, _mm256_set1_epi8(15));
	return _mm256_and_si256(_mm256_blendv_epi8(t0, t1, hi), valid);
}

// whole groups of 64 nucleotides; out 32-byte aligned when `stream`
__attribute__((target("avx2"))) void pack_avx2(const uint8_t *s, int64_t n64, uint8_t *out, bool stream)
{
	//                                      a  b   c  d   e  f  g  h   i  j  k   l  m  n   o  p
	const __m256i lut0 = _mm256_setr_epi8(1, 14, 2, 13, 0, 0, 4, 11, 0, 0, 12, 0, 3, 15, 0, 0,
					      1, 14, 2, 13, 0, 0, 4, 11, 0, 0, 12, 0, 3, 15, 0, 0);
	//                                      q  r  s  t  u  v  w  x  y   z
	const __m256i lut1 = _mm256_setr_epi8(0, 5, 6, 8, 8, 7, 9, 0, 10, 0, 0, 0, 0, 0, 0, 0,
					      0, 5, 6, 8, 8, 7, 9, 0, 10, 0, 0, 0, 0, 0, 0, 0);
	const __m256i mul = _mm256_set1_epi16(0x1001); // even character + 16 x odd character
	for (int64_t g = 0; g < n64; g++) {
		const __m256i a = codes_avx2(_mm256_loadu_si256((const __m256i *)(s + 64 * g)), lut0, lut1);
		const __m256i b = codes_avx2(_mm256_loadu_si256((const __m256i *)(s + 64 * g + 32)), lut0, lut1);
		const __m256i wa = _mm256_maddubs_epi16(a, mul), wb = _mm256_maddubs_epi16(b, mul);
		const __m256i r = _mm256_permute4x64_epi64(_mm256_packus_epi16(wa, wb), 0xd8);
		if (stream)
			_mm256_stream_si256((__m256i *)(out + 32 * g), r);
		else
			_mm256_storeu_si256((__m256i *)(out + 32 * g), r);
	}
	if (stream)
		_mm_sfence();
}
#endif

bool have_avx2()
{
#if defined(__x86_64__)
	static const bool yes = __builtin_cpu_supports("avx2") && getenv("GPUMOTIF_NO_AVX2") == NULL;
	return yes;
#else
	return false;
#endif
}

} // namespace

void host_pack_range(const uint8_t *s, int64_t n, uint8_t *out)
{
	int64_t done = 0;
#if defined(__x86_64__)
	if (have_avx2() && n >= 64) {
		const int64_t n64 = n / 64;
		pack_avx2(s, n64, out, ((uintptr_t)out & 31) == 0);
		done = n64 * 64;
	}
#endif
	if (done < n)
		pack_scalar(s + done, n - done, out + (done >> 1));
}

int host_pack_default_threads()
{
	const char *e = getenv("GPUMOTIF_PACK_THREADS");
	if (e != NULL && atoi(e) > 0)
		return std::min(64, atoi(e));
	int n = 0;
	cpu_set_t set;
	if (sched_getaffinity(0, sizeof set, &set) == 0)
		n = CPU_COUNT(&set);
	if (n <= 0)
		n = (int)std::thread::hardware_concurrency();
	return std::max(1, std::min(16, n));
}

// ------------------------------------------------------------------ thread team

PackTeam::PackTeam(int n_threads) : src_(NULL), dst_(NULL), n_(0), gen_(0), left_(0), quit_(false)
{
	n_threads = std::max(1, n_threads);
	for (int i = 1; i < n_threads; i++) // the caller of run() is member 0
		th_.emplace_back(&PackTeam::worker, this, i);
}

PackTeam::~PackTeam()
{
	{
		std::lock_guard<std::mutex> lk(m_);
		quit_ = true;
	}
	cv_work_.notify_all();
	for (auto &t : th_)
		t.join();
}

void PackTeam::piece(int k) const
{
	// pieces are multiples of 64 nucleotides (32 output bytes), the last takes the rest
	const int T = (int)th_.size() + 1;
	const int64_t per = ((n_ / T) + 63) & ~(int64_t)63;
	const int64_t lo = std::min<int64_t>(n_, per * k), hi = k == T - 1 ? n_ : std::min<int64_t>(n_, per * (k + 1));
	if (hi > lo)
		host_pack_range(src_ + lo, hi - lo, dst_ + (lo >> 1));
}

void PackTeam::worker(int k)
{
	int seen = 0;
	for (;;) {
		{
			std::unique_lock<std::mutex> lk(m_);
			cv_work_.wait(lk, [&] { return quit_ || gen_ != seen; });
			if (quit_)
				return;
			seen = gen_;
		}
		piece(k);
		{
			std::lock_guard<std::mutex> lk(m_);
			if (--left_ == 0)
				cv_done_.notify_all();
		}
	}
}

void PackTeam::run(const uint8_t *src, int64_t n, uint8_t *dst)
{
	{
		std::lock_guard<std::mutex> lk(m_);
		src_ = src;
		dst_ = dst;
		n_ = n;
		left_ = (int)th_.size();
		gen_++;
	}
	cv_work_.notify_all();
	piece(0);
	std::unique_lock<std::mutex> lk(m_);
	cv_done_.wait(lk, [&] { return left_ == 0; });
}

} // namespace gm

extern "C" int gm_host_pack(const char *seq, int64_t n, uint8_t *packed, int n_threads)
{
	if (n < 0 || (n > 0 && (seq == NULL || packed == NULL)))
		return -1;
	if (n_threads <= 0)
		n_threads = gm::host_pack_default_threads();
	if (n_threads == 1 || n < (1 << 16)) {
		gm::host_pack_range((const uint8_t *)seq, n, packed);
		return 0;
	}
	gm::PackTeam team(n_threads);
	team.run((const uint8_t *)seq, n, packed);
	return 0;
}
