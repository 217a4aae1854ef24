// TEST INFRASTRUCTURE ONLY -- not part of the shipped product path.
//
// Known-answer dump of the reference's unit functions on the hot path.  Every value
// written here is produced by the reference's own code (compiled in place from
// /root/reference/src): SignatureConstruction, RestWordBuffer/RestMatch, PopCount,
// AutoTextArray::getTextWord/isDontCareFree, RangeVector, ComputeScore and the Scoring
// table.  tests/golden/make_golden.py turns the JSON into committed fixtures.
#include "real_config.hpp"
#include "matchAllImplementation.cpp"
#include "harness_common.hpp"

#include <cstring>
#include <inttypes.h>

namespace
{
        struct Lcg
        {
                uint64_t s;
                Lcg(uint64_t seed) : s(seed) {}
                uint64_t next() { s = s * 6364136223846793005ULL + 1442695040888963407ULL; return s >> 17; }
        };

        uint32_t floatBits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
        uint64_t doubleBits(double f) { uint64_t u; memcpy(&u, &f, 8); return u; }

        template<typename signature_type>
        void dumpReadKats(FILE * out, RealOptions const & opts, std::vector<FASTQEntry> & reads, AutoTextArray<true> const & ATA, Scoring const & scoring)
        {
                SignatureConstruction<signature_type> const SC(opts.seedl, opts.nu);
                RestWordBuffer<true> RWB(opts.seedl);
                Lcg rng(12345);

                fprintf(out, "\"sig_shifts\":[%u,%u,%u,%u,%u,%u],\n", SC.s0shift(getSampleBits()), SC.s1shift(getSampleBits()), SC.s2shift(getSampleBits()),
                        SC.s3shift(getSampleBits()), SC.s4shift(getSampleBits()), SC.s5shift(getSampleBits()));
                fprintf(out, "\"reads\":[\n");
                for ( size_t r = 0; r < reads.size(); ++r )
                {
                        FASTQEntry const & P = reads[r];
                        unsigned int const patl = P.getPatternLength();
                        fprintf(out, "{\"patid\":%" PRIu64 ",\"len\":%u,\"mapped\":[", static_cast<uint64_t>(P.getPatID()), patl);
                        for ( unsigned int i = 0; i < patl; ++i ) fprintf(out, "%s%d", i?",":"", static_cast<int>(P.mapped[i]));
                        fprintf(out, "],\"transposed\":[");
                        for ( unsigned int i = 0; i < patl; ++i ) fprintf(out, "%s%d", i?",":"", static_cast<int>(P.transposed[i]));
                        fprintf(out, "],\"quality\":[");
                        for ( unsigned int i = 0; i < patl; ++i ) fprintf(out, "%s%u", i?",":"", P.getQuality(i));
                        fprintf(out, "]");

                        bool clean = patl >= static_cast<unsigned int>(opts.seedl);
                        for ( unsigned int i = 0; i < patl; ++i ) if ( P.mapped[i] > 3 ) clean = false;
                        fprintf(out, ",\"usable\":%d", clean ? 1 : 0);

                        if ( clean )
                        {
                                u_int32_t m[4], im[4] = {0,0,0,0};
                                SC.signatureMapped(P.mapped, &m[0]);
                                SC.reverseMappedSignature(P.mapped, &im[0]);
                                signature_type const fw[6] = { SC.s0(m[0],m[1]), SC.s1(m[0],m[2]), SC.s2(m[0],m[3]), SC.s3(m[1],m[2]), SC.s4(m[1],m[3]), SC.s5(m[2],m[3]) };
                                signature_type const rv[6] = { SC.s0(im[0],im[1]), SC.s1(im[0],im[2]), SC.s2(im[0],im[3]), SC.s3(im[1],im[2]), SC.s4(im[1],im[3]), SC.s5(im[2],im[3]) };
                                fprintf(out, ",\"m\":[%u,%u,%u,%u],\"im\":[%u,%u,%u,%u]", m[0],m[1],m[2],m[3], im[0],im[1],im[2],im[3]);
                                fprintf(out, ",\"fw\":[");
                                for ( int i = 0; i < 6; ++i ) fprintf(out, "%s%" PRIu64, i?",":"", static_cast<uint64_t>(fw[i]));
                                fprintf(out, "],\"rv\":[");
                                for ( int i = 0; i < 6; ++i ) fprintf(out, "%s%" PRIu64, i?",":"", static_cast<uint64_t>(rv[i]));
                                fprintf(out, "]");

                                RWB.setup(patl);
                                RWB.setupStraight(P.mapped);
                                RWB.setupReverse(P.mapped);
                                fprintf(out, ",\"fullrestwords\":%u,\"fracrestsyms\":%u,\"straightmatchoffset\":%u,\"reversematchoffset\":%u,\"straighttextrestoffset\":%d,\"reversetextrestoffset\":%d",
                                        RWB.fullrestwords, RWB.fracrestsyms, RWB.straightmatchoffset, RWB.reversematchoffset, RWB.straighttextrestoffset, RWB.reversetextrestoffset);
                                fprintf(out, ",\"rest_straight\":[");
                                for ( unsigned int i = 0; i < RWB.numrestwords; ++i ) fprintf(out, "%s%" PRIu64, i?",":"", static_cast<uint64_t>(RWB.Bstraight[i]));
                                fprintf(out, "],\"rest_reverse\":[");
                                for ( unsigned int i = 0; i < RWB.numrestwords; ++i ) fprintf(out, "%s%" PRIu64, i?",":"", static_cast<uint64_t>(RWB.Breverse[i]));
                                fprintf(out, "]");

                                // scores and rest distances of this read laid over a few text positions
                                fprintf(out, ",\"at\":[");
                                unsigned int emitted = 0;
                                for ( unsigned int q = 0; q < 6 && ATA.getN() >= patl; ++q )
                                {
                                        u_int64_t const pos = rng.next() % (ATA.getN() - patl + 1);
                                        if ( ! ATA.isDontCareFree(pos, patl) )
                                                continue;
                                        float const sf = ComputeScore<true,FASTQEntry,true>::computeScore(false, ATA, P, scoring, pos, patl);
                                        float const sr = ComputeScore<true,FASTQEntry,true>::computeScore(true, ATA, P, scoring, pos, patl);
                                        unsigned int const df = RestMatch<true>::computeDistance(RWB.Bstraight, RWB.fullrestwords, RWB.fracrestsyms, ATA, pos + RWB.straighttextrestoffset);
                                        unsigned int const dr = RestMatch<true>::computeDistance(RWB.Breverse, RWB.fullrestwords, RWB.fracrestsyms, ATA, pos);
                                        fprintf(out, "%s{\"pos\":%" PRIu64 ",\"score_fw\":%u,\"score_rv\":%u,\"rest_fw\":%u,\"rest_rv\":%u}", emitted?",":"",
                                                static_cast<uint64_t>(pos), floatBits(sf), floatBits(sr), df, dr);
                                        ++emitted;
                                }
                                fprintf(out, "]");
                        }
                        fprintf(out, "}%s\n", (r+1 < reads.size()) ? "," : "");
                }
                fprintf(out, "],\n");
        }
}

int harnessRunKat(RealOptions const & opts, std::string const & dumpname)
{
        FILE * out = fopen(dumpname.c_str(), "w");
        if ( ! out )
                throw std::runtime_error("cannot open dump file");

        std::vector< std::pair<std::string,u_int64_t> > ranges;
        std::auto_ptr< AutoTextArray<true> > AATA = getText<true>(opts.textfilename, ranges);
        AutoTextArray<true> const & ATA = *AATA;
        RangeVector<true> RV(ranges);
        Scoring const scoring(opts.similarity, opts.gc, opts.trans, opts.err, opts.gcmut_bias);

        int const qualityOffset = opts.qualityOffset ? opts.qualityOffset : 33;
        std::vector<FASTQEntry> reads;
        {
                u_int64_t const expect = FastQReader::countPatterns(opts.patternfilename);
                reads.resize(expect);
                FastQReader reader(opts.patternfilename, qualityOffset);
                u_int64_t got = 0;
                while ( got < expect && reader.getNextPatternUnlocked(reads[got]) ) { reads[got].computeMapped(); ++got; }
                reads.resize(got);
        }

        fprintf(out, "{\n\"seedl\":%d,\"n\":%" PRIu64 ",\n", opts.seedl, static_cast<uint64_t>(ATA.getN()));

        fprintf(out, "\"ranges\":[");
        for ( size_t i = 0; i < ranges.size(); ++i )
        {
                std::string esc;
                for ( size_t j = 0; j < ranges[i].first.size(); ++j )
                {
                        char const c = ranges[i].first[j];
                        if ( c == '"' || c == '\\' ) esc += '\\';
                        esc += c;
                }
                fprintf(out, "%s[\"%s\",%" PRIu64 "]", i?",":"", esc.c_str(), static_cast<uint64_t>(ranges[i].second));
        }
        fprintf(out, "],\n");

        // packed text words as the reference lays them out
        fprintf(out, "\"textwords\":[");
        u_int64_t const nwords = (ATA.getN()*2 + 63)/64;
        for ( u_int64_t i = 0; i < nwords; ++i )
                fprintf(out, "%s%" PRIu64, i?",":"", static_cast<uint64_t>(ATA.getTextWord(static_cast<unsigned int>(i))));
        fprintf(out, "],\n");

        fprintf(out, "\"symbols\":[");
        for ( u_int64_t i = 0; i < ATA.getN(); ++i )
                fprintf(out, "%s%u", i?",":"", static_cast<unsigned int>(ATA[i]));
        fprintf(out, "],\n");

        // unaligned extracts, wildcard and record predicates
        Lcg rng(777);
        fprintf(out, "\"textqueries\":[");
        for ( unsigned int q = 0; q < 400; ++q )
        {
                unsigned int const l = 1 + rng.next() % 32;
                if ( ATA.getN() < l ) continue;
                u_int64_t const i = rng.next() % (ATA.getN() - l + 1);
                unsigned int const patl = opts.seedl + rng.next() % 120;
                bool const dcf = ATA.isDontCareFree(i, l);
                bool const dcfp = (i + patl <= ATA.getN()) ? ATA.isDontCareFree(i, patl) : false;
                fprintf(out, "%s{\"i\":%" PRIu64 ",\"l\":%u,\"word\":%" PRIu64 ",\"dcf\":%d,\"patl\":%u,\"dcf_patl\":%d,\"inrange\":%d,\"valid\":%d,\"range\":%u}", q?",":"",
                        static_cast<uint64_t>(i), l, static_cast<uint64_t>(ATA.getTextWord(i,l)), dcf?1:0, patl, dcfp?1:0,
                        (i + patl <= ATA.getN()) ? 1 : 0,
                        RV.isPositionValid(i, patl) ? 1 : 0, RV.positionToRange(i));
        }
        fprintf(out, "],\n");

        fprintf(out, "\"diffcountpair64\":[");
        for ( unsigned int q = 0; q < 200; ++q )
        {
                u_int64_t a = (rng.next() << 40) ^ (rng.next() << 17) ^ rng.next();
                u_int64_t b = (q % 3 == 0) ? (a ^ (rng.next() & 0x3300c00f0ULL)) : ((rng.next() << 40) ^ (rng.next() << 17) ^ rng.next());
                fprintf(out, "%s[%" PRIu64 ",%" PRIu64 ",%u]", q?",":"", static_cast<uint64_t>(a), static_cast<uint64_t>(b), toollib::PopCount<true>::diffcountpair(a,b));
        }
        fprintf(out, "],\n\"diffcountpair32\":[");
        for ( unsigned int q = 0; q < 200; ++q )
        {
                u_int32_t a = static_cast<u_int32_t>(rng.next());
                u_int32_t b = (q % 3 == 0) ? (a ^ static_cast<u_int32_t>(rng.next() & 0x30c00f0U)) : static_cast<u_int32_t>(rng.next());
                fprintf(out, "%s[%u,%u,%u]", q?",":"", a, b, toollib::PopCount<true>::diffcountpair(a,b));
        }
        fprintf(out, "],\n");

        if ( opts.seedl <= 32 )
                dumpReadKats<u_int32_t>(out, opts, reads, ATA, scoring);
        else
                dumpReadKats<u_int64_t>(out, opts, reads, ATA, scoring);

        fprintf(out, "\"scoring_params\":[%.17g,%.17g,%.17g,%.17g,%.17g],\n", opts.similarity, opts.gc, opts.trans, opts.err, opts.gcmut_bias);
        fprintf(out, "\"filter_mult_bits\":%" PRIu64 ",\n", doubleBits(opts.filter_mult));
        fprintf(out, "\"ll_bits\":[");
        for ( unsigned int c0 = 0; c0 < 4; ++c0 )
                for ( unsigned int c1 = 0; c1 < 4; ++c1 )
                        for ( unsigned int q = 0; q < 64; ++q )
                                fprintf(out, "%s%" PRIu64, (c0|c1|q)?",":"", doubleBits(scoring.getRawLogScoreTable(c0,c1,q)));
        fprintf(out, "]\n}\n");
        fclose(out);
        return 0;
}
