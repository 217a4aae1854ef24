// TEST INFRASTRUCTURE ONLY -- not part of the shipped product path.
//
// matchAll leg of the reference harness.  This translation unit pulls in the reference's
// matchAllImplementation.cpp the same way its own instantiation units do
// (matchAllFastQ64sse4scores.cpp:19) and calls the reference's AllMatcher::match and
// unifyMatches unchanged.  What is written here is only a driver: it keeps all reads in
// memory, walks files and text blocks in the reference order
// (matchAllImplementation.cpp:403-476), and dumps every MatchPosAndError instead of
// formatting text lines (whose last <16 KiB per thread the stock writer drops).
#include "real_config.hpp"
#include "matchAllImplementation.cpp"
#include "harness_common.hpp"

#include <cstdlib>

namespace
{
        template<typename reader_type>
        void slurpReads(std::string const & filename, int const qualityOffset, std::vector<typename reader_type::pattern_type> & reads)
        {
                u_int64_t const expect = reader_type::countPatterns(filename);
                reads.resize(expect);
                reader_type reader(filename, qualityOffset);
                u_int64_t got = 0;
                // parse in place: Pattern keeps raw pointers into its own strings
                while ( got < expect && reader.getNextPatternUnlocked(reads[got]) )
                {
                        reads[got].computeMapped();
                        ++got;
                }
                reads.resize(got);
        }

        template<typename signature_type, typename reader_type, bool scores>
        int runAllTyped(RealOptions const & opts, std::string const & dumpname, HarnessTimes & times)
        {
                typedef typename reader_type::pattern_type pattern_type;
                typedef u_int32_t ptr_type;
                bool const sse4 = true;

                double const t_load0 = harnessNow();
                int const qualityOffset = opts.qualityOffset ? opts.qualityOffset : reader_type::getOffset(opts.patternfilename);
                if ( ! qualityOffset )
                        throw std::runtime_error("Unable to automatically detect FastQ quality format.");
                std::vector<pattern_type> reads;
                slurpReads<reader_type>(opts.patternfilename, qualityOffset, reads);
                times.reads = reads.size();
                times.load_s += harnessNow() - t_load0;

                std::pair<u_int64_t,u_int64_t> const read_info = getReadMemory<scores>(reads.size());
                std::vector<std::string> filenames;
                getFileList(opts.textfilename, filenames, ".fa");
                SignatureConstruction<signature_type> const SC(opts.seedl, opts.nu);
                Scoring const scoring(opts.similarity, opts.gc, opts.trans, opts.err, opts.gcmut_bias);

                FILE * dump = fopen(dumpname.c_str(), "wb");
                if ( ! dump )
                        throw std::runtime_error("cannot open dump file");

                char const * forced = getenv("REAL_HARNESS_NLIST");

                for ( unsigned int fi = 0; fi < filenames.size(); ++fi )
                {
                        double const t_text0 = harnessNow();
                        std::vector< std::pair<std::string,u_int64_t> > ranges;
                        std::auto_ptr< AutoTextArray<sse4> > AATA = getText<sse4>(filenames[fi], ranges);
                        AutoTextArray<sse4> const & ATA = *AATA;
                        RangeVector<sse4> RV(ranges);
                        times.load_s += harnessNow() - t_text0;
                        times.textlen += ATA.getN();

                        if ( ATA.getN() < static_cast<unsigned int>(opts.seedl) )
                                continue;

                        u_int64_t const filesize = ATA.getN();
                        u_int64_t n_list;
                        if ( forced )
                                n_list = strtoull(forced, 0, 10);
                        else
                        {
                                size_t const n_list_max = getNListMax<signature_type,ptr_type>(opts, read_info, ATA.size() + (ATA.getN()/8));
                                if ( ! n_list_max )
                                        throw std::bad_alloc();
                                u_int64_t const expblocks = (filesize - opts.seedl + 1 + (n_list_max-1)) / n_list_max;
                                n_list = (filesize + (expblocks-1)) / expblocks;
                        }

                        MapTextFile<signature_type,sse4> MTF(ATA, opts.seedl, opts.nu);
                        AllMatcher<signature_type,sse4,ptr_type,pattern_type,scores> AM(n_list, opts, SC, MTF, scoring, ATA, RV);

                        u_int32_t block = 0;
                        while ( true )
                        {
                                double const t_idx0 = harnessNow();
                                u_int64_t const masks = AM.readNextBlock();
                                times.index_s += harnessNow() - t_idx0;
                                if ( ! masks )
                                        break;
                                times.blocks += 1;

                                std::vector< std::vector<HarnessHit> > perthread;
                                double const t_m0 = harnessNow();
                                #if defined(_OPENMP)
                                #pragma omp parallel
                                #endif
                                {
                                        RestWordBuffer<sse4> RWB(opts.seedl);
                                        std::vector<HarnessHit> mine;
                                        u_int64_t handled = 0;

                                        #if defined(_OPENMP)
                                        #pragma omp for schedule(dynamic,4096)
                                        #endif
                                        for ( int64_t z = 0; z < static_cast<int64_t>(reads.size()); ++z )
                                        {
                                                std::vector<MatchPosAndError> found;
                                                AM.match(reads[z], fi, RWB, handled, found);
                                                if ( found.size() )
                                                {
                                                        unifyMatches(found);
                                                        for ( size_t q = 0; q < found.size(); ++q )
                                                        {
                                                                HarnessHit H;
                                                                H.patid = reads[z].getPatID();
                                                                H.pos = found[q].pos;
                                                                H.file = found[q].file;
                                                                H.frag = found[q].frag;
                                                                H.k = found[q].k;
                                                                H.inverted = found[q].inverted ? 1 : 0;
                                                                H.score = found[q].score;
                                                                H.block = block;
                                                                mine.push_back(H);
                                                        }
                                                }
                                        }

                                        #if defined(_OPENMP)
                                        #pragma omp critical
                                        #endif
                                        perthread.push_back(mine);
                                }
                                times.match_s += harnessNow() - t_m0;

                                for ( size_t t = 0; t < perthread.size(); ++t )
                                        if ( perthread[t].size() )
                                                fwrite(&perthread[t][0], sizeof(HarnessHit), perthread[t].size(), dump);
                                ++block;
                        }
                }

                fclose(dump);
                return 0;
        }

        template<typename signature_type, typename reader_type>
        int runAllScores(RealOptions const & opts, std::string const & dumpname, HarnessTimes & times)
        {
                if ( opts.scores )
                        return runAllTyped<signature_type,reader_type,true>(opts, dumpname, times);
                else
                        return runAllTyped<signature_type,reader_type,false>(opts, dumpname, times);
        }

        template<typename reader_type>
        int runAllWord(RealOptions const & opts, std::string const & dumpname, HarnessTimes & times)
        {
                // same word-size rule as real.cpp:217-220
                if ( opts.seedl <= 32 )
                        return runAllScores<u_int32_t,reader_type>(opts, dumpname, times);
                else
                        return runAllScores<u_int64_t,reader_type>(opts, dumpname, times);
        }
}

int harnessRunAll(RealOptions const & opts, std::string const & dumpname, HarnessTimes & times)
{
        // unlike real.cpp:325-328 the harness honours FastQ input in all-matches mode
        if ( opts.fastq )
                return runAllWord<FastQReader>(opts, dumpname, times);
        else
                return runAllWord<FastAReader>(opts, dumpname, times);
}
