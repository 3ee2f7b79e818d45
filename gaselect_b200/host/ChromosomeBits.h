// ChromosomeBits.h -- read access to the bit parts of the reference's Chromosome (src/Chromosome.h:78-85) without
// touching the reference's headers.
//
// The only public route from a Chromosome to its variables is toColumnSubset() (src/Chromosome.cpp:419-438), which walks
// the chromosome bit by bit: 1.8 us per call at p = 1000 -- 70 % of the host side of a generation once the fitness
// itself costs 4 us per chromosome.  The engine's C ABI takes the parts as they are (gsb_evaluate_bitsets: the layout of
// src/Chromosome.cpp:55-77,419-438), and the parts are also the exact cache key, so the batched driver never needs the
// column list at all.  `chromosomeParts` is private; the explicit-instantiation idiom below is standard C++ (access
// checks do not apply to the arguments of an explicit instantiation) and leaves src/Chromosome.h unmodified.  A
// maintainer merging the driver upstream would add `const std::vector<IntChromosome>& parts() const` instead
// (INTEGRATION.md).
#ifndef GASELECT_B200_CHROMOSOME_BITS_H
#define GASELECT_B200_CHROMOSOME_BITS_H

#include <cstdint>
#include <vector>

#include "Chromosome.h" // the reference's class, on the package's include path

namespace gsb_host {

template <typename Tag, typename Tag::type Member>
struct PrivateMember {
	friend typename Tag::type memberOf(Tag) { return Member; }
};
struct ChromosomePartsTag {
	typedef std::vector<IntChromosome> Chromosome::*type;
	friend type memberOf(ChromosomePartsTag);
};
template struct PrivateMember<ChromosomePartsTag, &Chromosome::chromosomeParts>;

static_assert(sizeof(IntChromosome) == sizeof(uint64_t), "gsb_evaluate_bitsets takes 64-bit parts");

// the parts of a chromosome: ceil(p / 64) words, variable v <-> bit (unusedBits + v) counted from part 0 (src/Chromosome.cpp:419-438)
inline const std::vector<IntChromosome>& partsOf(const Chromosome& ch) { return ch.*memberOf(ChromosomePartsTag()); }

} // namespace gsb_host

#endif
