// host_rng.hpp -- host-side generator and row shuffling, bit-exact with the reference.
//
// The CV segmentation is built on the host once per run and must be bit-identical to the
// reference's, so the generator is restated exactly:
//   * WELL19937a recurrence with the reference's tempering mask and index handling,
//     reference src/RNG.cpp:92-167 (six index cases folded into modular arithmetic here);
//   * seeding from one uint32 (src/RNG.cpp:75-90) and from a 624-word vector (:61-73), both
//     followed by a 500-draw burn-in (src/RNG.h:33);
//   * uniform(a, b) = a + u32 / 2^32 * (b - a) in double (src/RNG.h:21-23);
//   * ShuffledSet::shuffleAll on a persistent permutation (src/ShuffledSet.cpp:38-44).
#ifndef GASELECT_B200_HOST_RNG_HPP
#define GASELECT_B200_HOST_RNG_HPP

#include <cstdint>
#include <utility>
#include <vector>

namespace gsb {

class Well19937a {
public:
	static constexpr int kWords = 624;

	explicit Well19937a(uint32_t seed) {
		state_[0] = seed;
		for (uint32_t i = 1; i < kWords; ++i) {
			const uint32_t prev = state_[i - 1];
			state_[i] = 1812433253u * (prev ^ (prev >> 30)) + i;
		}
		pos_ = 0;
		burnIn();
	}

	explicit Well19937a(const uint32_t* seed624) {
		for (int i = 0; i < kWords; ++i) state_[i] = seed624[i];
		pos_ = 0;
		burnIn();
	}

	uint32_t next() {
		constexpr int M1 = 70, M2 = 179, M3 = 449;
		constexpr uint32_t upperBit = 0x80000000u, lowerBits = 0x7fffffffu, temper = 0x41180000u;
		const int i = pos_;
		// Reference quirk kept on purpose: for positions 0 and 1 the "previous" words are read at
		// i + M3 - 1 / i + M3 - 2, not at the wrapped positions (src/RNG.cpp:26,28 via :93,:106).
		const uint32_t back1 = (i >= 1) ? state_[i - 1] : state_[i + M3 - 1];
		const uint32_t back2 = (i >= 2) ? state_[i - 2] : state_[i + M3 - 2];
		const uint32_t a = state_[i];
		const uint32_t b = state_[wrap(i + M1)];
		const uint32_t c = state_[wrap(i + M2)];
		const uint32_t d = state_[wrap(i + M3)];
		const uint32_t z0 = (back1 & upperBit) | (back2 & lowerBits);
		const uint32_t z1 = (a ^ (a << 25)) ^ (b ^ (b >> 27));
		const uint32_t z2 = (c >> 9) ^ (d ^ (d >> 1));
		const uint32_t fresh = z1 ^ z2;
		const int j = wrap(i + kWords - 1);
		state_[i] = fresh;
		state_[j] = z0 ^ (z1 ^ (z1 << 9)) ^ (z2 ^ (z2 << 21)) ^ (fresh ^ (fresh >> 21));
		pos_ = j;
		return state_[j] ^ (state_[wrap(j + M2 + 1)] & temper);
	}

	double uniform(double lo, double hi) {
		return lo + (static_cast<double>(next()) / 4294967296.0) * (hi - lo);
	}

private:
	static int wrap(int i) { return (i >= kWords) ? i - kWords : i; }
	void burnIn() {
		for (int i = 0; i < 500; ++i) (void) next();
	}
	uint32_t state_[kWords];
	int pos_;
};

// Persistent permutation of 0..n-1; shuffleAll() continues from the previous permutation
// (the reference never resets the row set between replications, src/PLSEvaluator.cpp:98,145).
class RowShuffler {
public:
	explicit RowShuffler(uint32_t n) : set_(n) {
		for (uint32_t i = 0; i < n; ++i) set_[i] = i;
	}
	const std::vector<uint32_t>& shuffleAll(Well19937a& rng) {
		const size_t n = set_.size();
		for (size_t i = 0; i < n; ++i) {
			const size_t j = static_cast<size_t>(rng.uniform(static_cast<double>(i), static_cast<double>(n)));
			std::swap(set_[i], set_[j]);
		}
		return set_;
	}
private:
	std::vector<uint32_t> set_;
};

} // namespace gsb

#endif
