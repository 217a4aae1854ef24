// TEST INFRASTRUCTURE ONLY -- see bound_common.hpp.
//
// EnumerateAllMatches<...>::doMatching of the reference (matchAllImplementation.cpp:359-538) with its text-block loop
// (:451-535) replaced by real_gpu_match_all.  The rows come back as MatchPosAndError + the read ordinal, already in
// unifyMatches order; they are printed with the reference's own stream formatting of :485-510 (ids and bases from its
// reader, record names from its RangeVector).  Unlike the stock driver, which never flushes the tail of a block's
// output buffer (:512-517), every row is written.
#include "real_config.hpp"
#include "matchAllImplementation.cpp"         // the reference's types; its generic doMatching is never instantiated here
#include "bound_common.hpp"

namespace bound
{
        template<bool sse4, typename signature_type, typename reader_type, bool scores>
        void allDoMatching(RealOptions const & opts)
        {
                typedef typename reader_type::pattern_type pattern_type;
                unsigned int const numthreads = 1;
                std::cerr << "Starting matching for all hits on the GPU path (libreal_gpu.so)." << std::endl;
                u_int64_t const numpat = reader_type::countPatterns(opts.patternfilename);
                std::vector< std::string > filenames;
                getFileList(opts.textfilename, filenames, ".fa");
                int const qualityOffset = opts.qualityOffset ? opts.qualityOffset : reader_type::getOffset(opts.patternfilename);
                if ( ! qualityOffset )
                        throw std::runtime_error("Unable to automatically detect FastQ quality format.");
                Scoring const scoring(opts.similarity, opts.gc, opts.trans, opts.err, opts.gcmut_bias);

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
                std::vector<uint8_t> mapped, quality; std::vector<uint64_t> offsets;
                slurpReads<reader_type>(opts.patternfilename, qualityOffset, numthreads, mapped, quality, offsets, scores);
                if ( offsets.size() - 1 != numpat )
                        throw std::runtime_error("bound driver: pattern count mismatch");
                check(H.G, real_gpu_set_reads(H.G, mapped.empty() ? 0 : &mapped[0], quality.empty() ? 0 : &quality[0], &offsets[0], numpat), "set_reads");
                // the ids, in ordinal order, from the reference's id reader
                std::vector<std::string> ids;
                {
                        PatternIdReader<reader_type> pir(opts.patternfilename, qualityOffset, numthreads);
                        std::pair < typename reader_type::block_type *, typename reader_type::idblock_type * > block;
                        while ( pir.getBlock(block) )
                        {
                                for ( u_int64_t i = 0; i < block.first->blocksize; ++i ) ids.push_back(block.second->ids[i]);
                                pir.returnBlock(block);
                        }
                }

                std::auto_ptr < AsynchronousWriter  > output;
                if ( opts.outputfilename == "-" )
                        output = std::auto_ptr < AsynchronousWriter  >( new AsynchronousWriter(STDOUT_FILENO,16) );
                else
                        output = std::auto_ptr < AsynchronousWriter  >( new AsynchronousWriter(opts.outputfilename,16) );

                for ( unsigned int fi = 0; fi < filenames.size(); ++fi )
                {
                        bool const lastfile = (fi+1)==filenames.size();
                        std::cerr << "Processing file " << filenames[fi] << (lastfile?" (last processed file)":"")<< std::endl;
                        std::vector < std::pair < std::string, u_int64_t > > ranges;
                        std::auto_ptr < AutoTextArray<sse4> > AATA = getText<sse4>(filenames[fi],ranges);
                        AutoTextArray<sse4> const & ATA = *(AATA.get());
                        RangeVector<sse4> RV(ranges);
                        if ( ATA.getN() < static_cast<unsigned int>(opts.seedl) )
                        {
                                std::cerr << "File " << filenames[fi] << " is too small for seed length, skipping it." << std::endl;
                                continue;
                        }
                        u_int64_t const filesize = ATA.getN();
                        std::vector<uint64_t> words, nmask, starts;
                        textArrays(ATA, words, nmask);
                        for ( size_t i = 0; i < ranges.size(); ++i ) starts.push_back(ranges[i].second);
                        check(H.G, real_gpu_set_text(H.G, fi, &words[0], &nmask[0], filesize, 0, filesize, 0, filesize, &starts[0], (uint32_t)(starts.size() - 1)), "set_text");
                        real_gpu_hit const * hits = 0; uint64_t nhits = 0;
                        check(H.G, real_gpu_match_all(H.G, &hits, &nhits), "match_all");                 // replaces matchAllImplementation.cpp:451-535

                        std::ostringstream tempostr;
                        for ( uint64_t hi = 0; hi < nhits; ++hi )
                        {
                                real_gpu_hit const & M = hits[hi];
                                uint8_t const * m = &mapped[offsets[M.patid]];
                                unsigned int const patl = (unsigned int)(offsets[M.patid+1] - offsets[M.patid]);
                                std::string bases(reinterpret_cast<char const *>(m), reinterpret_cast<char const *>(m) + patl);
                                if ( M.inverted )
                                {
                                        // Pattern::transposed (Pattern.hpp:105-128): the reverse complement, wildcards kept
                                        std::string t(patl, 0);
                                        for ( unsigned int i = 0; i < patl; ++i ) { char const c = bases[patl-1-i]; t[i] = (c < 4) ? (3 - c) : c; }
                                        bases = t;
                                }
                                tempostr << ids[M.patid] << "\t" << toollib::remapString(bases) << "\t";
                                if ( scores )
                                        tempostr << M.score;
                                tempostr << "\t" << 1 << "\t" << "a" << "\t" << patl << "\t" << (M.inverted ? "-" : "+") << "\t"
                                         << RV.positionToId(M.pos) << "\t" << M.pos - ranges[RV.positionToRange(M.pos)].second + 1 << "\t" << "\t" << M.k << std::endl;
                                if ( tempostr.str().size() > 16384 )
                                {
                                        std::string const tempstring = tempostr.str();
                                        output->write ( tempstring.begin(), tempstring.end() );
                                        tempostr.str(std::string());
                                }
                        }
                        std::string const tempstring = tempostr.str();                          // the tail the stock driver drops
                        output->write ( tempstring.begin(), tempstring.end() );
                        std::cerr << "All done." << std::endl;
                }
        }
}

#define BOUND_ALL(S, W, Q) \
        template<> void EnumerateAllMatches<S, W, SLOW_ALL_FASTA_READER_TYPE, Q>::doMatching(RealOptions const & opts) { bound::allDoMatching<S, W, SLOW_ALL_FASTA_READER_TYPE, Q>(opts); }
#define BOUND_ALL_ALL(S) BOUND_ALL(S, u_int32_t, true) BOUND_ALL(S, u_int32_t, false) BOUND_ALL(S, u_int64_t, true) BOUND_ALL(S, u_int64_t, false)
BOUND_ALL_ALL(true)
BOUND_ALL_ALL(false)
