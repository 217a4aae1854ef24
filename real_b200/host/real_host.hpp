// Host side of the drop-in: the reference's run drivers re-hosted on the C ABI of libreal_gpu.so.
//
// Mirrors, with the same names and argument meaning,
//   RealOptions                         RealOptions.hpp:24-79, RealOptions.cpp:122-466
//   countLength / readFile / getText    countReads.cpp:28-125, getText.hpp:31-58
//   getFileList                         getFileList.cpp:155-174
//   FastAReader / FastQReader           FastAReader.hpp:107-138, FastQReader.hpp:130-180, 221-239
//   reorderFastA / reorderFastQ order   ReorderFastA.hpp:31-70, TemporaryFile.hpp:194-295 (-R 1)
//   EnumerateAllMatches::doMatching     matchAllImplementation.cpp:359-538
//   EnumerateUniqueMatches::doMatching  matchUniqueImplementation.cpp:1082-1489
//   Scoring                             Scoring.cpp:61-171
// The block loop over text-side index blocks (matchAllImplementation.cpp:451-535,
// matchUniqueImplementation.cpp:1253-1297) is what the library replaces: one real_gpu_set_reads
// (read-side index), then per text file one real_gpu_set_text + real_gpu_match_all / _match_unique.
// Nothing here matches on the CPU; library errors surface as std::runtime_error like the reference's.
#ifndef REAL_HOST_HPP
#define REAL_HOST_HPP

#include <stdint.h>
#include <memory>
#include <string>
#include <utility>
#include <vector>

namespace realhost
{

struct RealOptions
{
        static unsigned int const default_seedkmax = 2;
        static unsigned int const default_totalkmax = 5;
        static unsigned int const default_seedl = 32;
        static bool const default_match_unique = true;
        static bool const default_scores = true;
        static bool const default_rewritepatterns = true;
        static unsigned int const default_filter_level = 2;
        static unsigned int const nu = 4;

        std::string textfilename, patternfilename, outputfilename;
        unsigned int seedkmax, totalkmax;
        int seedl;
        bool match_unique;
        double fracmem;
        bool scores;
        unsigned int qualityOffset;
        bool rewritepatterns;
        int filter_level;
        double filter_mult;
        double similarity, err, trans, gc, gcmut_bias;
        bool gaps;
        bool fastq;
        int threads;            // -T: host threads that format the output (0 = all); the matching runs on the GPU
        int device;             // REAL_GPU_DEVICE, default 0
        int ngpus;              // REAL_GPUS, default 1: handles (= bucket shards) the matching is spread over, device + i modulo the
                                // devices present; the order dependent folds (-q 1, -g 1) always run on one handle
        bool rewritten_input;   // -p names a rewritten pattern file (first byte 0): its reads are in rewritten order already
        std::shared_ptr< std::vector<char> > stdin_bytes;   // -p -: the pattern file as read from standard input (RealOptions.cpp:418-426)

        RealOptions(int argc, char * argv[]);
        void printHelp() const;
        static bool isFastQ(std::string const & filename);
        double getFilterValue(unsigned int patl) const { return filter_mult * patl; }
};

// Scoring::LL, 4*4*64 doubles indexed (ref<<8)|(read<<6)|q
void buildScoringTable(double similarity, double gc, double trans, double err, double gcmut_bias, double * ll);

struct TextFile
{
        std::vector<uint64_t> words;     // 2 bit/base, MSB first (AutoTextArray.hpp:27-43), N stored as A
        std::vector<uint64_t> nmask;     // 1 bit/base, MSB first (AutoTextArray.hpp:45-61)
        uint64_t n;
        std::vector< std::pair<std::string, uint64_t> > ranges;   // (header text, start) + ("terminal", n)
        std::vector<uint64_t> starts() const;
};
void getText(std::string const & filename, TextFile & out);
void getFileList(std::string const & name, std::vector<std::string> & files, std::string const & suffix);

// the bytes of an input file, mapped read-only (or read into memory where the file cannot be mapped)
class FileBytes
{
        char const * p; size_t n; bool mapped; std::vector<char> owned;
        FileBytes(FileBytes const &); FileBytes & operator=(FileBytes const &);
        public:
        FileBytes() : p(0), n(0), mapped(false) {}
        ~FileBytes();
        void open(std::string const & filename);
        void adopt(std::vector<char> & bytes);          // takes the bytes over (standard input)
        void close();
        size_t size() const { return n; }
        char const * data() const { return p; }
        bool empty() const { return n == 0; }
        char const & operator[](size_t i) const { return p[i]; }
};

struct ReadSet
{
        std::vector<uint8_t> mapped;     // 0..3 = ACGT, 4 = anything else (Pattern.hpp:105-128)
        std::vector<uint8_t> quality;    // char - offset (FastQReader.hpp:168); empty for FASTA
        std::vector<uint64_t> offsets;   // nreads + 1
        std::vector<std::string> ids;
        uint64_t size() const { return offsets.size() - 1; }
};
int detectQualityOffset(std::string const & filename);                       // FastQReader::getOffset
void readPatterns(std::string const & filename, bool fastq, int qualityOffset, ReadSet & out, unsigned int threads = 0);   // 0 = all host threads
void readPatternsBuffer(FileBytes const & buf, bool fastq, int qualityOffset, ReadSet & out, unsigned int threads);
// the order the rewritten pattern file hands the reads out in (-R 1): by length, wildcard-free reads first
void reorderLikeRewrite(ReadSet & reads);
// The reference's rewritten pattern file itself (reorderFastA / reorderFastQ, ReorderFastA.hpp, ReorderFastQ.hpp,
// TemporaryFile.hpp:194-403; read back by FastDecoder.hpp:66-130, FastSubDecoder.hpp:53-168): per pattern length, ascending,
//     u32be length | ACGT section | ACGT ids | ACGTN section | ACGTN ids
// every section = u32be byte count (magic included) | u32be magic 0..3 | payload.  ACGT payload: per read ceil(L/4) bytes,
// 2 bit/base, first base in bits 7..6 (FASTQ: followed by the L quality values); ACGTN payload: per read ceil(L/2) bytes,
// 4 bit/base, first base in the high nibble (FASTQ: + L quality values); id payload: per read u16be length | bytes.
// The reference deletes the file when it is done (real.cpp:281-311); here it can be kept (REAL_KEEP_REWRITTEN=<path>) and
// given back as -p: a pattern file whose first byte is 0 is taken as a rewritten file, and its ACGT sections go to the
// device as they are.  A section of 4 GiB or more does not fit the format's 32-bit count (the reference writes the
// count modulo 2^32 and cannot read such a file back): it is written as count 0xFFFFFFFF followed by a u64be count.
// `reads` must be in rewritten order (reorderLikeRewrite).
void writeRewritten(ReadSet const & reads, bool fastq, std::vector<char> & out);
// the reference's memory planner (matchUniqueImplementation.cpp:1208-1244): seed windows per text-side index block
struct TextFile;
uint64_t planBlockWindows(RealOptions const & opts, TextFile const & T, uint64_t nreads);
bool looksRewritten(FileBytes const & buf);
// fills `reads` (in rewritten order) from the bytes of a rewritten file; fastq = whether its records carry qualities
void readRewritten(FileBytes const & buf, ReadSet & reads, bool & fastq);
bool rewrittenIsFastq(FileBytes const & buf);
// the same straight into the device's 2 bit/base layout (the ACGT sections copied as they stand), qualities and ids
struct PackedReads;
void readRewrittenPacked(FileBytes const & buf, PackedReads & out, std::vector<uint8_t> & quality, std::vector<char> & idbytes, std::vector<uint64_t> & idoff, bool & fastq);
// The reads 2 bit/base in the layout of the reference's rewritten pattern file (TemporaryFile.hpp:231-268,
// writePatternDontCareFree): 4 bases per byte, first base in bits 7..6, every read on a byte boundary; reads with a
// wildcard are flagged (the reference keeps them in a 4 bit/base section of their own) and stored as A.  This is what
// crosses PCIe (real_gpu_set_reads_packed): a quarter of the byte-per-base form.
struct PackedReads
{
        std::vector<uint8_t> packed;
        std::vector<uint64_t> byte_offsets;   // nreads + 1
        std::vector<uint32_t> lengths;
        std::vector<uint8_t> wildcard;
};
void packReads(ReadSet const & reads, PackedReads & out, unsigned int threads);

int doMatchingAll(RealOptions const & opts);
int doMatchingUnique(RealOptions const & opts);
int realMain(int argc, char * argv[]);

}
#endif
