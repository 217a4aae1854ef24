// TEST INFRASTRUCTURE ONLY -- not part of the shipped product path.
//
// matchUnique (+ optional gapped pass) leg of the reference harness.  Pulls in the
// reference's matchUniqueImplementation.cpp as its instantiation units do and calls
// UniqueMatcher::match / UniqueMatcher::matchGaps unchanged
// (matchUniqueImplementation.cpp:369-500, 501-572).  The driver below keeps the reads in
// memory, walks files and text blocks in the reference order
// (matchUniqueImplementation.cpp:1118-1297, 1302-1436) and dumps the raw
// UniqueMatchInfo words (+score) and the surviving GapInfo entries, which the stock CLI
// never prints.
#include "real_config.hpp"
#include "matchUniqueImplementation.cpp"
#include "harness_common.hpp"

#include <cstdlib>

namespace
{
        template<typename reader_type>
        void slurpReadsU(std::string const & filename, int const qualityOffset, std::vector<typename reader_type::pattern_type> & reads)
        {
                u_int64_t const expect = reader_type::countPatterns(filename);
                reads.resize(expect);
                reader_type reader(filename, qualityOffset);
                u_int64_t got = 0;
                while ( got < expect && reader.getNextPatternUnlocked(reads[got]) )
                {
                        reads[got].computeMapped();
                        ++got;
                }
                reads.resize(got);
        }

        // block length: the reference's memory planner (matchUniqueImplementation.cpp:1208-1244)
        // unless REAL_HARNESS_NLIST pins it
        template<typename signature_type, typename ptr_type, bool sse4>
        u_int64_t planBlock(RealOptions const & opts, AutoTextArray<sse4> const & ATA, RangeVector<sse4> const & RV, u_int64_t const infobytes)
        {
                char const * forced = getenv("REAL_HARNESS_NLIST");
                if ( forced )
                        return strtoull(forced, 0, 10);

                u_int64_t const fixed = ATA.size() + RV.size() + 2 * getNumLists() * getHistSize() * sizeof(size_t) + infobytes;
                if ( fixed > opts.usemem )
                        throw std::bad_alloc();
                u_int64_t const per = ((getNumLists()>>1) * sizeof(BaseMask<signature_type,ptr_type>) + ((getNumLists()>>1) + getRadixSortTemp()) * sizeof(Mask<signature_type,ptr_type>));
                u_int64_t const n_list_max = (opts.usemem - fixed) / per;
                if ( ! n_list_max )
                        throw std::bad_alloc();
                u_int64_t const filesize = ATA.getN();
                u_int64_t const expblocks = (filesize - opts.seedl + 1 + (n_list_max-1)) / n_list_max;
                return (filesize + (expblocks-1)) / expblocks;
        }

        template<typename signature_type, typename reader_type, bool scores>
        int runUniqueTyped(RealOptions const & opts, std::string const & dumpname, std::string const & gapdumpname, HarnessTimes & times)
        {
                typedef typename reader_type::pattern_type pattern_type;
                typedef u_int32_t ptr_type;
                bool const sse4 = true;

                double const t_load0 = harnessNow();
                int const qualityOffset = opts.qualityOffset ? opts.qualityOffset : reader_type::getOffset(opts.patternfilename);
                if ( ! qualityOffset )
                        throw std::runtime_error("Unable to automatically detect FastQ quality format.");
                std::vector<pattern_type> reads;
                slurpReadsU<reader_type>(opts.patternfilename, qualityOffset, reads);
                times.reads = reads.size();
                times.load_s += harnessNow() - t_load0;

                AutoArray< UniqueMatchInfo<scores> > uniqueinfo(reads.size());
                std::vector<std::string> filenames;
                getFileList(opts.textfilename, filenames, ".fa");
                SignatureConstruction<signature_type> const SC(opts.seedl, opts.nu);
                Scoring const scoring(opts.similarity, opts.gc, opts.trans, opts.err, opts.gcmut_bias);
                std::map<unsigned int, GapInfo> gapinfos;

                for ( int pass = 0; pass < (opts.gaps ? 2 : 1); ++pass )
                for ( unsigned int fi = 0; fi < filenames.size(); ++fi )
                {
                        double const t_text0 = harnessNow();
                        std::vector< std::pair<std::string,u_int64_t> > ranges;
                        std::auto_ptr< AutoTextArray<sse4> > AATA = getText<sse4>(filenames[fi], ranges);
                        AutoTextArray<sse4> const & ATA = *AATA;
                        RangeVector<sse4> RV(ranges);
                        times.load_s += harnessNow() - t_text0;
                        if ( pass == 0 )
                                times.textlen += ATA.getN();

                        if ( ATA.getN() < static_cast<unsigned int>(opts.seedl) )
                                continue;
                        if ( ranges.size() > UniqueMatchInfo<scores>::getMaxFragmentsPerFile() )
                                continue;

                        u_int64_t const n_list = planBlock<signature_type,ptr_type,sse4>(opts, ATA, RV, uniqueinfo.size());
                        MapTextFile<signature_type,sse4> MTF(ATA, opts.seedl, opts.nu);
                        UniqueMatcher<signature_type,sse4,ptr_type,pattern_type,scores> UM(n_list, opts, SC, MTF, scoring, ATA, RV);

                        while ( true )
                        {
                                double const t_idx0 = harnessNow();
                                u_int64_t const masks = UM.readNextBlock();
                                times.index_s += harnessNow() - t_idx0;
                                if ( ! masks )
                                        break;
                                times.blocks += 1;

                                double const t_m0 = harnessNow();
                                if ( pass == 0 )
                                {
                                        #if defined(_OPENMP)
                                        #pragma omp parallel
                                        #endif
                                        {
                                                RestWordBuffer<sse4> RWB(opts.seedl);
                                                u_int64_t handled = 0;
                                                #if defined(_OPENMP)
                                                #pragma omp for schedule(dynamic,4096)
                                                #endif
                                                for ( int64_t z = 0; z < static_cast<int64_t>(reads.size()); ++z )
                                                        UM.match(reads[z], uniqueinfo[reads[z].getPatID()], fi, RWB, handled);
                                        }
                                }
                                else
                                {
                                        // serial: the reference mutates the shared gapinfos map without a lock
                                        RestWordBuffer<sse4> RWB(opts.seedl);
                                        u_int64_t handled = 0;
                                        for ( size_t z = 0; z < reads.size(); ++z )
                                                UM.matchGaps(reads[z], uniqueinfo[reads[z].getPatID()], RWB, gapinfos, handled);
                                }
                                times.match_s += harnessNow() - t_m0;
                        }
                }

                FILE * dump = fopen(dumpname.c_str(), "wb");
                if ( ! dump )
                        throw std::runtime_error("cannot open dump file");
                for ( size_t i = 0; i < reads.size(); ++i )
                {
                        HarnessUnique U;
                        U.data = uniqueinfo[i].data;
                        U.score = scores ? uniqueinfo[i].getScore() : 0.0f;
                        U.pad = 0;
                        fwrite(&U, sizeof(U), 1, dump);
                }
                fclose(dump);

                if ( opts.gaps && gapdumpname.size() )
                {
                        FILE * gdump = fopen(gapdumpname.c_str(), "wb");
                        if ( ! gdump )
                                throw std::runtime_error("cannot open gap dump file");
                        for ( std::map<unsigned int,GapInfo>::const_iterator it = gapinfos.begin(); it != gapinfos.end(); ++it )
                        {
                                HarnessGap G;
                                G.patid = it->first;
                                G.mingap = it->second.MINgap;
                                G.where = it->second.where;
                                G.start = it->second.start;
                                G.gap_pos = it->second.gap_pos;
                                G.pad = 0;
                                fwrite(&G, sizeof(G), 1, gdump);
                        }
                        fclose(gdump);
                }

                return 0;
        }

        template<typename signature_type, typename reader_type>
        int runUniqueScores(RealOptions const & opts, std::string const & dumpname, std::string const & gapdumpname, HarnessTimes & times)
        {
                if ( opts.scores )
                        return runUniqueTyped<signature_type,reader_type,true>(opts, dumpname, gapdumpname, times);
                else
                        return runUniqueTyped<signature_type,reader_type,false>(opts, dumpname, gapdumpname, times);
        }

        template<typename reader_type>
        int runUniqueWord(RealOptions const & opts, std::string const & dumpname, std::string const & gapdumpname, HarnessTimes & times)
        {
                if ( opts.seedl <= 32 )
                        return runUniqueScores<u_int32_t,reader_type>(opts, dumpname, gapdumpname, times);
                else
                        return runUniqueScores<u_int64_t,reader_type>(opts, dumpname, gapdumpname, times);
        }
}

int harnessRunUnique(RealOptions const & opts, std::string const & dumpname, std::string const & gapdumpname, HarnessTimes & times)
{
        if ( opts.fastq )
                return runUniqueWord<FastQReader>(opts, dumpname, gapdumpname, times);
        else
                return runUniqueWord<FastAReader>(opts, dumpname, gapdumpname, times);
}
