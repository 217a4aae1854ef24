// Test helper: dumps what the host layer hands to the C ABI (packed text, parsed reads, scoring table, options)
// as JSON, so the CPU test-suite can check the parsers without a GPU.
#include "real_host.hpp"
#include "../csrc/fmt_g.h"
#include <cstdio>
#include <cmath>
#include <cstring>
#include <cstdlib>
#include <iostream>
#include <stdexcept>

using namespace realhost;

static void jstr(std::string const & s)
{
        putchar('"');
        for ( size_t i = 0; i < s.size(); ++i )
        {
                unsigned char const c = s[i];
                if ( c == '"' || c == '\\' ) { putchar('\\'); putchar(c); }
                else if ( c < 32 ) printf("\\u%04x", c);
                else putchar(c);
        }
        putchar('"');
}

int main(int argc, char * argv[])
{
        try
        {
                if ( argc < 2 ) return 2;
                std::string const mode = argv[1];
                if ( mode == "text" )
                {
                        TextFile T; getText(argv[2], T);
                        printf("{\"n\":%llu,\"ranges\":[", (unsigned long long)T.n);
                        for ( size_t i = 0; i < T.ranges.size(); ++i ) { if ( i ) putchar(','); putchar('['); jstr(T.ranges[i].first); printf(",%llu]", (unsigned long long)T.ranges[i].second); }
                        printf("],\"words\":[");
                        for ( size_t i = 0; i < T.words.size(); ++i ) printf("%s%llu", i ? "," : "", (unsigned long long)T.words[i]);
                        printf("],\"nmask\":[");
                        for ( size_t i = 0; i < T.nmask.size(); ++i ) printf("%s%llu", i ? "," : "", (unsigned long long)T.nmask[i]);
                        printf("]}\n");
                }
                else if ( mode == "reads" )
                {
                        bool const fastq = atoi(argv[3]); int qoff = atoi(argv[4]); bool const rewrite = atoi(argv[5]);
                        if ( fastq && ! qoff ) qoff = detectQualityOffset(argv[2]);
                        ReadSet R; readPatterns(argv[2], fastq, qoff, R);
                        if ( rewrite ) reorderLikeRewrite(R);
                        printf("{\"qoff\":%d,\"ids\":[", qoff);
                        for ( size_t i = 0; i < R.ids.size(); ++i ) { if ( i ) putchar(','); jstr(R.ids[i]); }
                        printf("],\"offsets\":[");
                        for ( size_t i = 0; i < R.offsets.size(); ++i ) printf("%s%llu", i ? "," : "", (unsigned long long)R.offsets[i]);
                        printf("],\"mapped\":[");
                        for ( size_t i = 0; i < R.mapped.size(); ++i ) printf("%s%u", i ? "," : "", (unsigned)R.mapped[i]);
                        printf("],\"quality\":[");
                        for ( size_t i = 0; i < R.quality.size(); ++i ) printf("%s%u", i ? "," : "", (unsigned)R.quality[i]);
                        printf("]}\n");
                }
                else if ( mode == "rewrite" )
                {
                        // <pattern file> <fastq> <quality offset> <out>: the reference's rewritten pattern file of it (writeRewritten)
                        bool const fastq = atoi(argv[3]); int qoff = atoi(argv[4]);
                        if ( fastq && ! qoff ) qoff = detectQualityOffset(argv[2]);
                        ReadSet R; readPatterns(argv[2], fastq, qoff, R);
                        reorderLikeRewrite(R);
                        std::vector<char> bytes; writeRewritten(R, fastq, bytes);
                        FILE * f = fopen(argv[5], "wb");
                        if ( ! f || (bytes.size() && fwrite(&bytes[0], 1, bytes.size(), f) != bytes.size()) ) throw std::runtime_error("cannot write");
                        fclose(f);
                        printf("{\"bytes\":%llu}\n", (unsigned long long)bytes.size());
                }
                else if ( mode == "unrewrite" )
                {
                        FileBytes buf; buf.open(argv[2]);
                        ReadSet R; bool fastq = false;
                        if ( ! looksRewritten(buf) ) throw std::runtime_error("not a rewritten pattern file");
                        readRewritten(buf, R, fastq);
                        printf("{\"fastq\":%d,\"ids\":[", (int)fastq);
                        for ( size_t i = 0; i < R.ids.size(); ++i ) { if ( i ) putchar(','); jstr(R.ids[i]); }
                        printf("],\"offsets\":[");
                        for ( size_t i = 0; i < R.offsets.size(); ++i ) printf("%s%llu", i ? "," : "", (unsigned long long)R.offsets[i]);
                        printf("],\"mapped\":[");
                        for ( size_t i = 0; i < R.mapped.size(); ++i ) printf("%s%u", i ? "," : "", (unsigned)R.mapped[i]);
                        printf("],\"quality\":[");
                        for ( size_t i = 0; i < R.quality.size(); ++i ) printf("%s%u", i ? "," : "", (unsigned)R.quality[i]);
                        printf("]}\n");
                }
                else if ( mode == "plan" )
                {
                        // <text bases> <reads> -- <REAL options>: the n_list the memory planner chooses (its messages go to stderr)
                        TextFile T; T.n = strtoull(argv[2], 0, 10);
                        uint64_t const nreads = strtoull(argv[3], 0, 10);
                        RealOptions o(argc - 3, argv + 3);
                        printf("{\"n_list\":%llu}\n", (unsigned long long)planBlockWindows(o, T, nreads));
                }
                else if ( mode == "unrewrite_packed" )
                {
                        // the rewritten file straight into the device layout (readRewrittenPacked)
                        FileBytes buf; buf.open(argv[2]);
                        PackedReads P; std::vector<uint8_t> q; std::vector<char> idb; std::vector<uint64_t> ido; bool fastq = false;
                        readRewrittenPacked(buf, P, q, idb, ido, fastq);
                        printf("{\"fastq\":%d,\"lengths\":[", (int)fastq);
                        for ( size_t i = 0; i < P.lengths.size(); ++i ) printf("%s%u", i ? "," : "", P.lengths[i]);
                        printf("],\"wildcard\":[");
                        for ( size_t i = 0; i < P.wildcard.size(); ++i ) printf("%s%u", i ? "," : "", (unsigned)P.wildcard[i]);
                        printf("],\"byte_offsets\":[");
                        for ( size_t i = 0; i < P.byte_offsets.size(); ++i ) printf("%s%llu", i ? "," : "", (unsigned long long)P.byte_offsets[i]);
                        printf("],\"packed\":[");
                        for ( size_t i = 0; i < (size_t)P.byte_offsets.back(); ++i ) printf("%s%u", i ? "," : "", (unsigned)P.packed[i]);
                        printf("],\"quality\":[");
                        for ( size_t i = 0; i < q.size(); ++i ) printf("%s%u", i ? "," : "", (unsigned)q[i]);
                        printf("],\"ids\":[");
                        for ( size_t i = 0; i + 1 < ido.size(); ++i ) { if ( i ) putchar(','); jstr(std::string(&idb[ido[i]], &idb[ido[i+1]])); }
                        printf("]}\n");
                }
                else if ( mode == "ll" )
                {
                        double ll[1024];
                        buildScoringTable(atof(argv[2]), atof(argv[3]), atof(argv[4]), atof(argv[5]), atof(argv[6]), ll);
                        printf("[");
                        for ( int i = 0; i < 1024; ++i ) { unsigned long long u; memcpy(&u, &ll[i], 8); printf("%s%llu", i ? "," : "", u); }
                        printf("]\n");
                }
                else if ( mode == "opts" )
                {
                        RealOptions o(argc - 1, argv + 1);
                        unsigned long long fm; memcpy(&fm, &o.filter_mult, 8);
                        printf("{\"seedkmax\":%u,\"totalkmax\":%u,\"seedl\":%d,\"match_unique\":%d,\"scores\":%d,\"qualityOffset\":%u,\"rewritepatterns\":%d,\"filter_level\":%d,"
                               "\"filter_mult_bits\":%llu,\"gaps\":%d,\"fastq\":%d}\n", o.seedkmax, o.totalkmax, o.seedl, (int)o.match_unique, (int)o.scores, o.qualityOffset,
                               (int)o.rewritepatterns, o.filter_level, fm, (int)o.gaps, (int)o.fastq);
                }
                else if ( mode == "files" )
                {
                        std::vector<std::string> f; getFileList(argv[2], f, ".fa");
                        printf("[");
                        for ( size_t i = 0; i < f.size(); ++i ) { if ( i ) putchar(','); jstr(f[i]); }
                        printf("]\n");
                }
                else if ( mode == "fmtg" )
                {
                        // the score formatter of the device (csrc/fmt_g.h) against snprintf("%g") on `count` floats: pseudo-random bit
                        // patterns, values near the powers of ten (rounding up into the next exponent), halves, small integers
                        unsigned long long const count = strtoull(argv[2], 0, 10);
                        unsigned long long x = 0x9E3779B97F4A7C15ULL, bad = 0, done = 0;
                        auto check = [&](float f)
                        {
                                char a[32], b[32];
                                int const na = fmtg::format_g6(f, a);
                                int const nb = snprintf(b, sizeof(b), "%g", (double)f);
                                ++done;
                                if ( na != nb || memcmp(a, b, na) )
                                {
                                        if ( f != f && na == nb - 0 && (nb == 3 || nb == 4) ) { }       // nan: compared below by its letters only
                                        a[na] = 0;
                                        if ( ! (f != f) && bad++ < 10 ) { unsigned int u; memcpy(&u, &f, 4); fprintf(stderr, "%08x: ours '%s' libc '%s'\n", u, a, b); }
                                }
                        };
                        for ( unsigned long long i = 0; i < count; ++i )
                        {
                                x ^= x << 13; x ^= x >> 7; x ^= x << 17;
                                unsigned int u = (unsigned int)(x >> 16);
                                float f; memcpy(&f, &u, 4);
                                check(f);
                                // scores live here: a few hundred at most, six digits cut in the middle of the mantissa
                                check((float)((double)(long long)(x % 2000000) / 997.0 - 1000.0));
                        }
                        for ( int e = -45; e <= 38; ++e )
                                for ( int k = -40; k <= 40; ++k )
                                {
                                        double const v = pow(10.0, e);
                                        float f = (float)v;
                                        for ( int s = 0; s < (k < 0 ? -k : k); ++s ) f = nextafterf(f, k < 0 ? 0.0f : INFINITY);
                                        check(f); check(-f);
                                        check((float)(v * 9.999995)); check((float)(v * 9.9999949)); check((float)(v * 1.2345650)); check((float)(v * 1.2345649));
                                }
                        for ( int i = 0; i < 3000000; ++i ) { check((float)i * 0.5f); check((float)i * 0.03125f); }
                        check(0.0f); check(-0.0f); check(INFINITY); check(-INFINITY);
                        printf("{\"checked\":%llu,\"bad\":%llu}\n", done, bad);
                }
                else return 2;
                return 0;
        }
        catch ( std::exception const & e ) { std::cerr << e.what() << std::endl; return 1; }
}
