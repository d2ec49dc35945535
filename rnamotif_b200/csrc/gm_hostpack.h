// gm_hostpack.h -- host-side 4-bit packer (gm_hostpack.cpp)
#pragma once
#include <cstdint>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

namespace gm {

// n characters at s -> (n + 1) / 2 bytes at out, on the calling thread
void host_pack_range(const uint8_t *s, int64_t n, uint8_t *out);
// GPUMOTIF_PACK_THREADS, else the CPUs this process may run on, at most 16
int host_pack_default_threads();

// persistent team: run() packs one chunk with every member and returns when it is done
class PackTeam {
public:
	explicit PackTeam(int n_threads);
	~PackTeam();
	void run(const uint8_t *src, int64_t n, uint8_t *dst);
	int size() const { return (int)th_.size() + 1; }

private:
	void worker(int k);
	void piece(int k) const;
	std::vector<std::thread> th_;
	std::mutex m_;
	std::condition_variable cv_work_, cv_done_;
	const uint8_t *src_;
	uint8_t *dst_;
	int64_t n_;
	int gen_, left_;
	bool quit_;
};

} // namespace gm

extern "C" int gm_host_pack(const char *seq, int64_t n, uint8_t *packed, int n_threads);
