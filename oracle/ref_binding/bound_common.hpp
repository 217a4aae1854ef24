// TEST INFRASTRUCTURE ONLY -- the reference-side binding of INTEGRATION.md section 3, compiled for real.
//
// oracle/Makefile target `ref_bound` builds oracle/_ref/real_bound from
//   * the reference's OWN real.cpp, RealOptions.cpp, countReads.cpp, Scoring.cpp ... compiled where they lie under
//     /root/reference/src (nothing is copied), and
//   * bound_unique.cpp / bound_all.cpp (this directory), which take the place of the reference's 64 instantiation units
//     matchUniqueFast*.cpp / matchAllFast*.cpp: they include the reference's implementation files for its types (readers,
//     getText, AutoTextArray, RangeSet, UniqueMatchInfo, printMatchUnlocked, Scoring ...) and SPECIALISE
//     EnumerateUniqueMatches<...>::doMatching / EnumerateAllMatches<...>::doMatching -- the symbols real.cpp calls
//     (real.cpp:203-212) -- with drivers whose text-block loop is the C ABI of include/real_gpu.h.
// So `real_bound` is the reference binary (its option parser, its pattern rewriting, its readers, its text loader, its
// output code) with the matching replaced by libreal_gpu.so; tests/test_ref_binding_gpu.py compares its output files with
// the stock binary's (tests/golden/cli_*.txt).
#ifndef BOUND_COMMON_HPP
#define BOUND_COMMON_HPP

#include "real_gpu.h"

#include <stdexcept>
#include <string>
#include <vector>
#include <cstring>

namespace bound
{
        inline void check(real_gpu * G, int rc, char const * what)
        {
                if ( rc != REAL_GPU_OK )
                        throw std::runtime_error(std::string(what) + ": " + (G ? real_gpu_last_error(G) : "library error"));
        }

        struct Handle
        {
                real_gpu * G;
                Handle() : G(0) {}
                ~Handle() { if ( G ) real_gpu_destroy(G); }
        };

        // the scoring table through the reference's own accessor, so that libm never enters (Scoring.hpp:70-73)
        template<typename scoring_type>
        void scoringTable(scoring_type const & scoring, double * ll)
        {
                for ( unsigned int c0 = 0; c0 < 4; ++c0 )
                        for ( unsigned int c1 = 0; c1 < 4; ++c1 )
                                for ( unsigned int q = 0; q < 64; ++q )
                                        ll[(c0<<8)|(c1<<6)|q] = scoring.getRawLogScoreTable(c0,c1,q);
        }

        // all reads once, as Pattern::computeMapped leaves them (Pattern.hpp:105-128): mapped bytes, qualities, offsets.
        // The read ordinal (Pattern::patid) is the position in the file the reader hands out -- the rewritten file with -R 1.
        template<typename reader_type>
        void slurpReads(std::string const & filename, int const qualityOffset, unsigned int const numthreads,
                        std::vector<uint8_t> & mapped, std::vector<uint8_t> & quality, std::vector<uint64_t> & offsets, bool const wantquality)
        {
                typedef typename reader_type::pattern_type pattern_type;
                reader_type patfile(filename, qualityOffset);
                typename reader_type::stream_data_type ad(patfile, FastFileDecoderBase::default_blocksize, std::max(4u,3*numthreads));
                typename reader_type::stream_reader_type ar(ad);
                typename reader_type::block_type * block = 0;
                offsets.assign(1, 0);
                u_int64_t expect = 0;
                while ( (block = ar.getBlock()) )
                {
                        for ( u_int64_t z = 0; z < block->blocksize; ++z )
                        {
                                pattern_type const & pattern = block->getPattern(z);
                                if ( pattern.getPatID() != expect++ )
                                        throw std::runtime_error("bound driver: the reader does not hand the reads out in ordinal order");
                                unsigned int const patl = pattern.getPatternLength();
                                mapped.insert(mapped.end(), pattern.mapped, pattern.mapped + patl);
                                if ( wantquality )
                                        for ( unsigned int i = 0; i < patl; ++i )
                                                quality.push_back(static_cast<uint8_t>(pattern.getQuality(i)));
                                offsets.push_back(mapped.size());
                        }
                        ar.returnBlock(block);
                }
        }

        // AutoTextArray -> the two arrays of real_gpu_set_text.  The packed words are public (getTextWord(i)); the wildcard
        // words have no accessor in the reference, so the mask is rebuilt base by base through operator[] -- in a maintained
        // patch this is a one-line accessor to AutoTextArray::wild (AutoTextArray.hpp:18-24).
        template<typename ata_type>
        void textArrays(ata_type const & ATA, std::vector<uint64_t> & words, std::vector<uint64_t> & nmask)
        {
                u_int64_t const n = ATA.getN();
                words.assign((n + 31) / 32 + 1, 0);
                nmask.assign((n + 63) / 64 + 1, 0);
                for ( u_int64_t w = 0; w < (n + 31) / 32; ++w )
                        words[w] = ATA.getTextWord(static_cast<unsigned int>(w));
                for ( u_int64_t i = 0; i < n; ++i )
                        if ( ATA[i] & 4 )
                                nmask[i >> 6] |= 1ULL << (63 - (i & 63));
        }
}
#endif
