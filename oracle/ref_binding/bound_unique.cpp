// TEST INFRASTRUCTURE ONLY -- see bound_common.hpp.
//
// EnumerateUniqueMatches<...>::doMatching of the reference (matchUniqueImplementation.cpp:1082-1489) with its text-block
// loop (:1253-1297) and its gapped second pass (:1302-1436) replaced by the C ABI.  Everything around the loop is the
// reference's own code, called, not restated: countPatterns, getFileList, Scoring, getText (AutoTextArray), RangeSet,
// the reader of the (rewritten) pattern file, printMatchUnlocked and the AsynchronousWriter of the final pass.
#include "real_config.hpp"
#include "matchUniqueImplementation.cpp"      // the reference's types; its generic doMatching is never instantiated here
#include "bound_common.hpp"

namespace bound
{
        // the reference's memory planner (matchUniqueImplementation.cpp:1208-1244): windows per text-side index block.  Only the
        // order dependent folds read it (real_gpu_set_block_windows); REAL_NLIST pins it like in bin/real.
        template<typename signature_type, typename ptr_type, bool sse4>
        u_int64_t planBlockWindows(RealOptions const & opts, AutoTextArray<sse4> const & ATA, RangeVector<sse4> const & RV, u_int64_t const infobytes)
        {
                if ( char const * forced = getenv("REAL_NLIST") )
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

        template<bool sse4, typename signature_type, typename reader_type, bool scores>
        void uniqueDoMatching(RealOptions const & opts)
        {
                typedef typename reader_type::pattern_type pattern_type;
                unsigned int const numthreads = 1;
                std::cerr << "Starting matching for unique best hits on the GPU path (libreal_gpu.so)." << std::endl;

                u_int64_t const numpat = reader_type::countPatterns(opts.patternfilename);
                AutoArray < UniqueMatchInfo<scores> > uniqueinfo(numpat);
                std::vector< std::string > filenames;
                getFileList(opts.textfilename, filenames, ".fa");
                u_int64_t unique = 0;
                int const qualityOffset = opts.qualityOffset ? opts.qualityOffset : reader_type::getOffset(opts.patternfilename);
                if ( ! qualityOffset )
                        throw std::runtime_error("Unable to automatically detect FastQ quality format.");
                Scoring const scoring(opts.similarity, opts.gc, opts.trans, opts.err, opts.gcmut_bias);
                RangeSet RS;

                // ---- the handle and the read-side index (replaces the construction of SignatureConstruction / UniqueMatcher per file)
                Handle H;
                real_gpu_params P; memset(&P, 0, sizeof(P));
                P.struct_size = sizeof(P); P.device = 0;
                if ( char const * e = getenv("REAL_GPU_DEVICE") ) P.device = atoi(e);
                P.seedl = opts.seedl; P.seedkmax = opts.seedkmax; P.totalkmax = opts.totalkmax; P.scores = scores ? 1 : 0;
                P.filter_mult = opts.filter_mult;
                double ll[1024];
                scoringTable(scoring, ll);
                P.ll_table = ll;
                if ( real_gpu_create(&P, &H.G) != REAL_GPU_OK )
                        throw std::runtime_error("real_gpu_create failed (no CUDA device or unsupported option); there is no CPU fallback");
                {
                        std::vector<uint8_t> mapped, quality; std::vector<uint64_t> offsets;
                        slurpReads<reader_type>(opts.patternfilename, qualityOffset, numthreads, mapped, quality, offsets, scores || opts.gaps);
                        if ( offsets.size() - 1 != numpat )
                                throw std::runtime_error("bound driver: pattern count mismatch");
                        check(H.G, real_gpu_set_reads(H.G, mapped.empty() ? 0 : &mapped[0], quality.empty() ? 0 : &quality[0], &offsets[0], numpat), "set_reads");
                }

                std::vector<bool> fileused(filenames.size(), false);
                for ( int pass = 0; pass < (opts.gaps ? 2 : 1); ++pass )
                for ( unsigned int fi = 0; fi < filenames.size(); ++fi )
                {
                        bool const lastfile = (fi+1)==filenames.size();
                        std::cerr << "Processing file " << filenames[fi] << (lastfile?" (last processed file)":"")<< std::endl;
                        std::vector < std::pair < std::string, u_int64_t > > ranges;
                        std::auto_ptr < AutoTextArray<sse4> > AATA = getText<sse4>(filenames[fi],ranges);            // the reference's loader
                        AutoTextArray<sse4> const & ATA = *(AATA.get());
                        RangeVector<sse4> RV(ranges);
                        if ( pass == 0 )
                                RS.addRange(ranges);
                        u_int64_t const filesize = ATA.getN();
                        std::cerr << "File size = " << filesize << std::endl;
                        if ( ATA.getN() < static_cast<unsigned int>(opts.seedl) )
                        {
                                std::cerr << "File " << filenames[fi] << " is too small for seed length, skipping it." << std::endl;
                                continue;
                        }
                        if ( ranges.size() > UniqueMatchInfo<scores>::getMaxFragmentsPerFile() )
                        {
                                std::cerr << "Number of fragments " << ranges.size() << " in file is larger than limit " << UniqueMatchInfo<scores>::getMaxFragmentsPerFile() << " we can handle, skipping it." << std::endl;
                                continue;
                        }
                        std::vector<uint64_t> words, nmask, starts;
                        textArrays(ATA, words, nmask);
                        for ( size_t i = 0; i < ranges.size(); ++i ) starts.push_back(ranges[i].second);
                        check(H.G, real_gpu_set_text(H.G, fi, &words[0], &nmask[0], filesize, 0, filesize, 0, filesize, &starts[0], (uint32_t)(starts.size() - 1)), "set_text");
                        u_int64_t const n_list = planBlockWindows<signature_type,u_int32_t,sse4>(opts, ATA, RV, uniqueinfo.size());
                        std::cerr << "Using n_list = " << n_list << std::endl;
                        if ( pass == 0 )
                        {
                                check(H.G, real_gpu_set_block_windows(H.G, n_list), "set_block_windows");
                                check(H.G, real_gpu_match_unique(H.G), "match_unique");          // replaces matchUniqueImplementation.cpp:1253-1297
                        }
                        else
                                check(H.G, real_gpu_match_gaps(H.G, n_list), "match_gaps");        // replaces :1302-1436
                        std::cerr << "All done." << std::endl;
                }
                {
                        std::vector<uint64_t> words(numpat + 1); std::vector<float> sc(numpat + 1);
                        check(H.G, real_gpu_get_unique(H.G, &words[0], scores ? &sc[0] : 0), "get_unique");
                        for ( u_int64_t i = 0; i < numpat; ++i )
                        {
                                uniqueinfo[i].data = words[i];
                                uniqueinfo[i].setScore(sc[i]);                         // (a no-op without scores, UniqueMatchInfo.hpp:178)
                        }
                }

                // ---- the reference's final pass, unchanged (matchUniqueImplementation.cpp:1438-1486)
                PatternIdReader<reader_type> pir (opts.patternfilename, qualityOffset, numthreads);
                std::pair < typename reader_type::block_type *, typename reader_type::idblock_type * > block;
                std::auto_ptr < AsynchronousWriter  > aw;
                if ( opts.outputfilename == "-" )
                        aw = std::auto_ptr < AsynchronousWriter  >( new AsynchronousWriter(STDOUT_FILENO,16) );
                else
                        aw = std::auto_ptr < AsynchronousWriter  >( new AsynchronousWriter(opts.outputfilename,16) );
                while (  pir.getBlock(block) )
                {
                        for ( u_int64_t i = 0; i < block.first->blocksize; ++i )
                        {
                                pattern_type const & pattern = block.first->getPattern(i);
                                std::string const & id = block.second->ids[i];
                                UniqueMatchInfo<scores> & info = uniqueinfo[pattern.getPatID()];
                                std::ostringstream ostr;
                                printMatchUnlocked<scores>(info,ostr,pattern,id,pattern.getPatternLength(),unique,RS);
                                std::string const os = ostr.str();
                                aw->write(os.begin(), os.end());
                        }
                        pir.returnBlock(block);
                }
                std::cerr << "unique: " << unique << std::endl;
        }
}

// the symbols real.cpp links against (real.cpp:203-212 dispatches on SSE4, word size, reader and scores)
#define BOUND_UNIQUE(S, W, R, Q) \
        template<> void EnumerateUniqueMatches<S, W, R, Q>::doMatching(RealOptions const & opts) { bound::uniqueDoMatching<S, W, R, Q>(opts); }
#define BOUND_UNIQUE_READERS(S, W, Q) \
        BOUND_UNIQUE(S, W, SLOW_UNIQUE_FASTA_READER_TYPE, Q) BOUND_UNIQUE(S, W, SLOW_UNIQUE_FASTQ_READER_TYPE, Q) \
        BOUND_UNIQUE(S, W, FAST_UNIQUE_FASTA_READER_TYPE, Q) BOUND_UNIQUE(S, W, FAST_UNIQUE_FASTQ_READER_TYPE, Q)
#define BOUND_UNIQUE_ALL(S) \
        BOUND_UNIQUE_READERS(S, u_int32_t, true) BOUND_UNIQUE_READERS(S, u_int32_t, false) \
        BOUND_UNIQUE_READERS(S, u_int64_t, true) BOUND_UNIQUE_READERS(S, u_int64_t, false)
BOUND_UNIQUE_ALL(true)
BOUND_UNIQUE_ALL(false)
