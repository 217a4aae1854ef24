// TEST INFRASTRUCTURE ONLY -- not part of the shipped product path.
//
// Shared declarations of the reference harness (oracle/_ref/ref_harness).  The harness is
// linked against the reference's OWN translation units, compiled in place from
// /root/reference/src by oracle/Makefile.  It re-states nothing of the matching
// algorithm: it only drives AllMatcher::match / UniqueMatcher::match /
// UniqueMatcher::matchGaps (matchAllImplementation.cpp:261, matchUniqueImplementation.cpp:369,501)
// over reads held in memory and dumps their results in binary, because the stock CLI
// cannot emit complete matchAll / gapped results.
#ifndef ORACLE_REF_HARNESS_COMMON_HPP
#define ORACLE_REF_HARNESS_COMMON_HPP

#include <string>
#include <vector>
#include <cstdio>
#include <stdint.h>
#include <sys/time.h>

struct RealOptions;

// One matchAll result row as dumped to disk (little endian, 40 bytes).
struct HarnessHit
{
        uint64_t patid;
        uint64_t pos;
        uint32_t file;
        uint32_t frag;
        uint32_t k;
        uint32_t inverted;
        float score;
        uint32_t block;
};

// One matchUnique result row (16 bytes): the packed UniqueMatchInfo word and the score
// (0 when the run was made without scores).
struct HarnessUnique
{
        uint64_t data;
        float score;
        uint32_t pad;
};

// One surviving GapInfo map entry of the gapped pass (24 bytes).
struct HarnessGap
{
        uint32_t patid;
        uint32_t mingap;
        uint32_t where;
        uint32_t start;
        uint32_t gap_pos;
        uint32_t pad;
};

struct HarnessTimes
{
        double load_s;   // text load + read parse
        double index_s;  // readNextBlock(): MapTextFile + radix sorts + lookup tables
        double match_s;  // OpenMP matching region
        uint64_t blocks;
        uint64_t reads;
        uint64_t textlen;
        HarnessTimes() : load_s(0), index_s(0), match_s(0), blocks(0), reads(0), textlen(0) {}
};

inline double harnessNow()
{
        struct timeval tv;
        gettimeofday(&tv, 0);
        return tv.tv_sec + 1e-6 * tv.tv_usec;
}

int harnessRunAll(RealOptions const & opts, std::string const & dumpname, HarnessTimes & times);
int harnessRunUnique(RealOptions const & opts, std::string const & dumpname, std::string const & gapdumpname, HarnessTimes & times);
int harnessRunKat(RealOptions const & opts, std::string const & dumpname);

#endif
