// See real_host.hpp.  Host C++ above the C ABI (include/real_gpu.h).
#include "real_host.hpp"
#include "../../include/real_gpu.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <iterator>
#include <sstream>
#include <stdexcept>
#include <thread>
#include <chrono>

#include <dirent.h>
#include <sys/stat.h>
#include <unistd.h>
#include <fcntl.h>
#include <sys/mman.h>

namespace realhost
{

// ---------------------------------------------------------------------------------------------
// RealOptions (RealOptions.cpp:85-466)
// ---------------------------------------------------------------------------------------------

static double const DFLT_SIMILARITY = 0.995, DFLT_GC = 0.41, DFLT_TRANS = 0.71, DFLT_ERR = 0.00, DFLT_GCMUT_BIAS = 2.0;   // Scoring.cpp:204-208
static double const default_fracmem = 0.75;

void RealOptions::printHelp() const
{
        std::cerr << "Options:" << std::endl;
        std::cerr << "-t <textfilename>" << std::endl;
        std::cerr << "-p <patternfilename>" << std::endl;
        std::cerr << "-o <outputfilename>" << std::endl;
        std::cerr << "-s <maximum number of errors in seed, default=" << default_seedkmax << ">" << std::endl;
        std::cerr << "-e <total maximum number of errors, default=" << default_totalkmax << ">" << std::endl;
        std::cerr << "-l <length of seed, default=" << default_seedl << ">" << std::endl;
        std::cerr << "-u <search for unique match, default=" << default_match_unique << ">" << std::endl;
        std::cerr << "-f <fraction of physical memory to use, default=" << default_fracmem << ">" << std::endl;
        std::cerr << "-q <use quality scores, default=" << default_scores << ">" << std::endl;
        std::cerr << "-Q <offset for quality scores, default=autodetect>" << std::endl;
        std::cerr << "-R <rewrite pattern file, default=" << default_rewritepatterns << ">" << std::endl;
        std::cerr << "-T <number of host threads (output formatting); the matching runs on the GPU>" << std::endl;
        std::cerr << "-similarity <sequence similarity, default=" << DFLT_SIMILARITY << ">" << std::endl;
        std::cerr << "-trans <transitions fraction of mutations, default=" << DFLT_TRANS << ">" << std::endl;
        std::cerr << "-gc <composition bias, default=" << DFLT_GC << ">" << std::endl;
        std::cerr << "-gcmut_bias <mutability bias of G&C, default=" << DFLT_GCMUT_BIAS << ">" << std::endl;
        std::cerr << "-filter_level <filtering level for equal hits 0-4, default=" << default_filter_level << ">" << std::endl;
}

bool RealOptions::isFastQ(std::string const & filename)
{
        std::ifstream istr(filename.c_str());
        if ( ! istr.is_open() )
                throw std::runtime_error("Unable to open pattern file.");
        int const first = istr.get();
        if ( first < 0 )
                throw std::runtime_error("Failed to read first character from pattern file.");
        if ( first == '>' ) return false;
        if ( first == '@' ) return true;
        if ( first == 0 ) return false;         // a rewritten pattern file kept from an earlier run (readRewritten decides)
        throw std::runtime_error("Unable to determine type of pattern file.");
}

static double clampWarn(char const * name, double v, bool upper)
{
        if ( v < 0 ) { std::cerr << "Warning: setting " << name << " up to 0." << std::endl; v = 0; }
        if ( upper && v > 1 ) { std::cerr << "Warning: setting " << name << " down to 1." << std::endl; v = 1; }
        return v;
}

RealOptions::RealOptions(int argc, char * argv[])
: seedkmax(default_seedkmax), totalkmax(default_totalkmax), seedl(default_seedl), match_unique(default_match_unique), fracmem(default_fracmem),
  scores(default_scores), qualityOffset(0), rewritepatterns(default_rewritepatterns), filter_level(default_filter_level), filter_mult(0),
  similarity(DFLT_SIMILARITY), err(DFLT_ERR), trans(DFLT_TRANS), gc(DFLT_GC), gcmut_bias(DFLT_GCMUT_BIAS), gaps(false), fastq(false), threads(0), device(0)
{
        std::vector<std::string> opts;
        for ( int i = 1; i < argc; ++i )
                opts.push_back(argv[i]);
        unsigned int i = 0;
        while ( i < opts.size() )
        {
                std::string const & a = opts[i];
                bool const known2 = a == "-t" || a == "-p" || a == "-o" || a == "-s" || a == "-e" || a == "-l" || a == "-u" || a == "-g" || a == "-R" || a == "-m" ||
                        a == "-q" || a == "-Q" || a == "-f" || a == "-T" || a == "-similarity" || a == "-err" || a == "-trans" || a == "-gc" || a == "-gcmut_bias" || a == "-filter_level";
                if ( known2 )
                {
                        if ( i + 1 >= opts.size() )
                                throw std::runtime_error("Parameter for argument " + a + " is missing.");
                        char const * v = opts[i+1].c_str();
                        if ( a == "-t" ) textfilename = v;
                        else if ( a == "-p" ) patternfilename = v;
                        else if ( a == "-o" ) outputfilename = v;
                        else if ( a == "-s" ) seedkmax = atoi(v);
                        else if ( a == "-e" )
                        {
                                totalkmax = atoi(v);
                                if ( totalkmax > 15 )                   // UniqueMatchInfoBase::getMaxErrors()
                                {
                                        totalkmax = 15;
                                        std::cerr << "Warning: reducing maximum amount of errors to " << totalkmax << std::endl;
                                }
                        }
                        else if ( a == "-l" ) seedl = atoi(v);
                        else if ( a == "-u" ) match_unique = atoi(v);
                        else if ( a == "-g" ) gaps = atoi(v);
                        else if ( a == "-R" ) rewritepatterns = atoi(v);
                        else if ( a == "-m" || a == "-f" ) fracmem = atof(v);
                        else if ( a == "-q" ) scores = atoi(v);
                        else if ( a == "-Q" ) qualityOffset = atoi(v);
                        else if ( a == "-T" )
                        {
                                threads = atoi(v);
                                if ( threads < 1 )
                                        throw std::runtime_error("Argument for -T parameter is invalid (<1)");
                        }
                        else if ( a == "-similarity" )
                        {
                                similarity = clampWarn("similarity", atof(v), true);
                                if ( similarity == 0 ) std::cerr << "Warning: similarity value is " << similarity << std::endl;
                        }
                        else if ( a == "-err" )
                        {
                                err = clampWarn("err", atof(v), true);
                                if ( err == 1 ) std::cerr << "Warning: err value is " << err << std::endl;
                        }
                        else if ( a == "-trans" )
                        {
                                trans = clampWarn("trans", atof(v), true);
                                if ( trans == 0 || trans == 1 ) std::cerr << "Warning: trans value is " << trans << std::endl;
                        }
                        else if ( a == "-gc" )
                        {
                                gc = clampWarn("gc", atof(v), true);
                                if ( gc == 0 || gc == 1 ) std::cerr << "Warning: gc value is " << gc << std::endl;
                        }
                        else if ( a == "-gcmut_bias" )
                        {
                                gcmut_bias = clampWarn("gcmut_bias", atof(v), false);
                                if ( gcmut_bias == 0 ) std::cerr << "Warning: gcmut_bias value is " << gcmut_bias << std::endl;
                        }
                        else if ( a == "-filter_level" )
                        {
                                filter_level = atoi(v);
                                if ( filter_level < 0 ) { std::cerr << "Warning: setting filter_level up to 0." << std::endl; filter_level = 0; }
                                if ( filter_level > 4 ) { std::cerr << "Warning: setting filter_level down to 4." << std::endl; filter_level = 4; }
                        }
                        i += 2;
                }
                else if ( a == "-h" )
                {
                        printHelp();
                        throw std::runtime_error("Help requested.");
                }
                else
                {
                        std::cerr << "Ignoring argument " << a << std::endl;
                        i += 1;
                }
        }
        if ( ! (textfilename.size() && patternfilename.size() && outputfilename.size()) )
                printHelp();
        if ( ! textfilename.size() )
                throw std::runtime_error("Mandatory argument -t (text file name) is not given.");
        if ( ! patternfilename.size() )
                throw std::runtime_error("Mandatory argument -p (pattern file name) is not given.");
        if ( ! outputfilename.size() )
                throw std::runtime_error("Mandatory argument -o (output file name) is not given.");
        fracmem = std::min(1.0, fracmem);
        rewritten_input = false;
        ngpus = 1;
        if ( char const * e = getenv("REAL_GPUS") ) ngpus = std::max(1, atoi(e));
        if ( patternfilename == "-" )
        {
                // RealOptions.cpp:418-426: the type is decided by the first byte of standard input, rewriting is switched on.
                // The whole stream is taken in here (the reference spools it into its rewritten pattern file).
                stdin_bytes.reset(new std::vector<char>());
                char tmp[1 << 16];
                ssize_t got;
                while ( (got = ::read(STDIN_FILENO, tmp, sizeof(tmp))) > 0 ) stdin_bytes->insert(stdin_bytes->end(), tmp, tmp + got);
                if ( stdin_bytes->empty() )
                        throw std::runtime_error("Failed to read first character from pattern file.");
                char const first = (*stdin_bytes)[0];
                if ( first == '>' ) fastq = false;
                else if ( first == '@' ) fastq = true;
                else throw std::runtime_error("Unable to determine type of pattern file.");
                if ( ! rewritepatterns )
                {
                        std::cerr << "Reading patterns from stdin, switching on pattern rewriting." << std::endl;
                        rewritepatterns = true;
                }
        }
        else
        {
                fastq = isFastQ(patternfilename);
                FileBytes head;
                head.open(patternfilename);
                if ( looksRewritten(head) )
                {
                        // the reference's rewritten pattern file, kept from an earlier run (REAL_KEEP_REWRITTEN): the reads come out of
                        // it in rewritten order, qualities already reduced by their offset
                        rewritten_input = true;
                        fastq = rewrittenIsFastq(head);
                        std::cerr << "pattern file is a rewritten pattern file" << std::endl;
                }
        }
        std::cerr << "pattern file is " << (fastq ? "FASTQ" : "FASTA") << " rewrite is " << (rewritepatterns ? "on" : "off") << std::endl;
        if ( seedl > 64 )
        {
                seedl = 64;
                std::cerr << "reduced seed size to " << seedl << " to not exceed 64." << std::endl;
        }
        if ( seedl % 4 )
        {
                seedl -= (seedl % 4);
                std::cerr << "reduced seed size to " << seedl << " to have a multiple of 4." << std::endl;
        }
        if ( seedl < static_cast<int>(nu) )
                throw std::runtime_error("cannot handle seed length < 4");
        if ( seedkmax > 2 )
        {
                seedkmax = 2;
                std::cerr << "reduced number of mismatches in seed to " << seedkmax << " as we cannot handle more." << std::endl;
        }
        switch ( filter_level )
        {
                case 1: filter_mult = 0.5 * totalkmax; break;
                case 2: filter_mult = 1 * totalkmax; break;
                case 3: filter_mult = 2 * totalkmax; break;
                case 4: filter_mult = 3 * totalkmax; break;
                case 0: default: filter_mult = 0 * totalkmax; break;
        }
        filter_mult /= 70.0;
        std::cerr << "filter_mult=" << filter_mult << std::endl;
        if ( char const * e = getenv("REAL_GPU_DEVICE") ) device = atoi(e);
}

// ---------------------------------------------------------------------------------------------
// Scoring table (Scoring.cpp:61-133 init, :155-171 getScore)
// ---------------------------------------------------------------------------------------------

static double const q_prb[65] = {
        1.0000000, 0.7943282, 0.6309573, 0.5011872, 0.3981072, 0.3162278, 0.2511886, 0.1995262, 0.1584893, 0.1258925,
        0.1000000, 0.0794328, 0.0630957, 0.0501187, 0.0398107, 0.0316228, 0.0251189, 0.0199526, 0.0158489, 0.0125893,
        0.0100000, 0.0079433, 0.0063096, 0.0050119, 0.0039811, 0.0031623, 0.0025119, 0.0019953, 0.0015849, 0.0012589,
        0.0010000, 0.0007943, 0.0006310, 0.0005012, 0.0003981, 0.0003162, 0.0002512, 0.0001995, 0.0001585, 0.0001259,
        0.0001000, 0.0000794, 0.0000631, 0.0000501, 0.0000398, 0.0000316, 0.0000251, 0.0000200, 0.0000158, 0.0000126,
        0.0000100, 0.0000079, 0.0000063, 0.0000050, 0.0000040, 0.0000032, 0.0000025, 0.0000020, 0.0000016, 0.0000013,
        0.0000010, 0.0000008, 0.0000006, 0.0000005, 0.0000004 };

void buildScoringTable(double similarity, double gc, double trans, double err, double gcmut_bias, double * ll)
{
        volatile double odds[4][4];     // volatile: every intermediate is rounded to double like the reference's -ffloat-store build
        double const transit = trans * (1 - similarity);
        double const transver = (1 - trans) * (1 - similarity);
        double const bg[4] = { (1 - gc) / 2, gc / 2, gc / 2, (1 - gc) / 2 };
        double const bias = gcmut_bias * (1 - gc) / gc;
        odds[0][2] = transit / (bias + 1) / (1 - gc);
        odds[3][1] = transit / (bias + 1) / (1 - gc);
        odds[2][0] = transit / (bias + 1) / gc * bias;
        odds[1][3] = transit / (bias + 1) / gc * bias;
        odds[0][1] = transver / 2 / (bias + 1) / (1 - gc);
        odds[3][2] = transver / 2 / (bias + 1) / (1 - gc);
        odds[0][3] = transver / 2 / (bias + 1) / (1 - gc);
        odds[3][0] = transver / 2 / (bias + 1) / (1 - gc);
        odds[1][0] = transver / 2 / (bias + 1) / gc * bias;
        odds[2][3] = transver / 2 / (bias + 1) / gc * bias;
        odds[1][2] = transver / 2 / (bias + 1) / gc * bias;
        odds[2][1] = transver / 2 / (bias + 1) / gc * bias;
        odds[0][0] = 1 - odds[0][1] - odds[0][2] - odds[0][3];
        odds[3][3] = 1 - odds[3][0] - odds[3][1] - odds[3][2];
        odds[2][2] = 1 - odds[2][0] - odds[2][1] - odds[2][3];
        odds[1][1] = 1 - odds[1][0] - odds[1][2] - odds[1][3];
        for ( int x = 0; x < 4; ++x )
                for ( int y = 0; y < 4; ++y )
                {
                        odds[x][y] = odds[x][y] * (1 - err);
                        odds[x][y] = odds[x][y] / bg[y];
                }
        double const log2 = std::log(2.0);
        for ( unsigned int c0 = 0; c0 < 4; ++c0 )
                for ( unsigned int c1 = 0; c1 < 4; ++c1 )
                        for ( unsigned int q = 0; q < 64; ++q )
                        {
                                volatile double const l = std::log(odds[c0][c1]) / log2;
                                ll[(c0 << 8) | (c1 << 6) | q] = l * (1 - q_prb[q]);
                        }
}

// ---------------------------------------------------------------------------------------------
// text (countReads.cpp:28-125, AutoTextArray.hpp:27-61)
// ---------------------------------------------------------------------------------------------

// the bytes of a file: mapped read-only (no copy, the page cache is the buffer), or read into memory where mapping fails
FileBytes::~FileBytes() { close(); }

void FileBytes::close()
{
        if ( mapped && p ) munmap(const_cast<char *>(p), n);
        p = 0; n = 0; mapped = false;
        std::vector<char>().swap(owned);
}

void FileBytes::open(std::string const & filename)
{
        close();
        int const fd = ::open(filename.c_str(), O_RDONLY);
        if ( fd < 0 )
                throw std::runtime_error("Failed to open file " + filename);
        struct stat st;
        if ( fstat(fd, &st) == 0 && S_ISREG(st.st_mode) && st.st_size > 0 )
        {
                void * m = mmap(0, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
                if ( m != MAP_FAILED )
                {
                        madvise(m, (size_t)st.st_size, MADV_WILLNEED);
                        p = static_cast<char const *>(m); n = (size_t)st.st_size; mapped = true;
                        ::close(fd);
                        return;
                }
        }
        // not mappable (or empty): read it
        char tmp[1 << 16];
        ssize_t got;
        while ( (got = ::read(fd, tmp, sizeof(tmp))) > 0 ) owned.insert(owned.end(), tmp, tmp + got);
        ::close(fd);
        if ( got < 0 )
                throw std::runtime_error("Failed to read file " + filename);
        p = owned.empty() ? 0 : &owned[0]; n = owned.size();
}

void FileBytes::adopt(std::vector<char> & bytes)
{
        close();
        owned.swap(bytes);
        p = owned.empty() ? 0 : &owned[0]; n = owned.size();
}

static void slurp(std::string const & filename, FileBytes & buf) { buf.open(filename); }

std::vector<uint64_t> TextFile::starts() const
{
        std::vector<uint64_t> s(ranges.size());
        for ( size_t i = 0; i < ranges.size(); ++i ) s[i] = ranges[i].second;
        return s;
}

void getText(std::string const & filename, TextFile & out)
{
        std::cerr << "Computing length of file " << filename << "...";
        FileBytes buf;
        slurp(filename, buf);
        out.ranges.clear();
        out.words.assign(buf.size() / 32 + 2, 0);
        out.nmask.assign(buf.size() / 64 + 2, 0);
        // one pass with the state machine of countLength/readFile: '>' starts a header that runs to the end of
        // the line, everything but A C G T N outside headers is dropped (lower case included)
        bool header = false;
        uint64_t cnt = 0, idcnt = 0;
        std::string id;
        for ( size_t p = 0; p < buf.size(); ++p )
        {
                char const c = buf[p];
                if ( c == '>' ) { header = true; idcnt = cnt; id.resize(0); }
                else if ( c == '\n' )
                {
                        if ( header )
                                out.ranges.push_back(std::pair<std::string, uint64_t>(id, idcnt));
                        header = false;
                }
                else if ( header ) id += c;
                else
                {
                        unsigned int code;
                        switch ( c )
                        {
                                case 'A': code = 0; break;
                                case 'C': code = 1; break;
                                case 'G': code = 2; break;
                                case 'T': code = 3; break;
                                case 'N': code = 4; break;
                                default: continue;
                        }
                        if ( code == 4 )
                                out.nmask[cnt >> 6] |= 1ULL << (63 - (cnt & 63));
                        else
                                out.words[cnt >> 5] |= (uint64_t)code << (62 - 2 * (cnt & 31));
                        ++cnt;
                }
        }
        out.ranges.push_back(std::pair<std::string, uint64_t>("terminal", cnt));
        out.n = cnt;
        out.words.resize((cnt + 31) / 32 + 1);
        out.nmask.resize((cnt + 63) / 64 + 1);
        std::cerr << "done, length is " << cnt << std::endl;
}

static bool endsOn(std::string const & s, std::string const & suffix)
{
        return s.size() >= suffix.size() && s.compare(s.size() - suffix.size(), suffix.size(), suffix) == 0;
}

static void enumerateFilesInDirectory(std::string const & dirname, std::vector<std::string> & files, std::string const & suffix)
{
        DIR * d = opendir(dirname.c_str());
        if ( ! d )
                throw std::runtime_error("Could not open file/directory");
        struct dirent * e = 0;
        while ( (e = readdir(d)) )
        {
                std::string const filename = dirname + "/" + e->d_name;
                struct stat st;
                if ( stat(filename.c_str(), &st) == 0 )
                {
                        if ( S_ISREG(st.st_mode) )
                        {
                                if ( endsOn(filename, suffix) )
                                        files.push_back(filename);
                        }
                        else if ( S_ISDIR(st.st_mode) && std::string(e->d_name) != "." && std::string(e->d_name) != ".." )
                                enumerateFilesInDirectory(filename + "/", files, suffix);
                }
        }
        closedir(d);
}

void getFileList(std::string const & name, std::vector<std::string> & files, std::string const & suffix)
{
        struct stat st;
        if ( stat(name.c_str(), &st) != 0 )
                return;
        if ( S_ISREG(st.st_mode) && endsOn(name, suffix) )
                files.push_back(name);
        else if ( S_ISDIR(st.st_mode) )
                enumerateFilesInDirectory(name, files, suffix);
}

// ---------------------------------------------------------------------------------------------
// reads (FastAReader.hpp:107-138, FastQReader.hpp:130-180)
// ---------------------------------------------------------------------------------------------

namespace
{
        struct Cursor
        {
                FileBytes const & b; size_t p;
                Cursor(FileBytes const & rb) : b(rb), p(0) {}
                int get() { return p < b.size() ? (unsigned char)b[p++] : -1; }
        };
        inline uint8_t mapChar(int c)
        {
                switch ( c ) { case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3; default: return 4; }
        }
}

// the reader state machines, one record at a time; quality == 0 for FASTA.  Same transitions as the reference's
// character-at-a-time loops (FastAReader.hpp:107-138, FastQReader.hpp:130-180), taken a run of bytes at a time.
static inline bool isSpaceByte(unsigned char c) { return c == ' ' || (c >= 9 && c <= 13); }      // isspace in the "C" locale

static bool nextPattern(Cursor & in, bool fastq, int qualityOffset, bool & foundnextmarker, std::string & id, std::string & pat, std::string * quality)
{
        char const marker = fastq ? '@' : '>';
        if ( ! foundnextmarker )
                return false;
        foundnextmarker = false;
        char const * const B = in.b.data();
        size_t const N = in.b.size();
        size_t p = in.p;
        {
                // id: the rest of the marker's line
                void const * e = p < N ? memchr(B + p, '\n', N - p) : 0;
                if ( ! e ) { in.p = N; return false; }
                size_t const q = static_cast<char const *>(e) - B;
                id.assign(B + p, q - p);
                p = q + 1;
        }
        pat.resize(0);
        unsigned char const patterm = fastq ? '+' : '>';
        int c = -1;
        while ( p < N )
        {
                unsigned char const ch = B[p];
                if ( ch == patterm ) { c = ch; ++p; break; }
                if ( isSpaceByte(ch) ) { ++p; continue; }
                size_t q = p + 1;
                while ( q < N && (unsigned char)B[q] != patterm && ! isSpaceByte(B[q]) ) ++q;
                pat.append(B + p, q - p);
                p = q;
        }
        if ( ! fastq )
        {
                in.p = p;
                foundnextmarker = (c == '>');
                return true;
        }
        {
                // rest of the '+' line
                void const * e = p < N ? memchr(B + p, '\n', N - p) : 0;
                if ( ! e ) { in.p = N; return false; }
                p = static_cast<char const *>(e) - B + 1;
        }
        quality->resize(0);
        // the reference's loop fetches one character beyond the last quality value (FastQReader.hpp:163-165)
        for ( ;; )
        {
                if ( p >= N ) break;
                unsigned char const ch = B[p++];
                if ( quality->size() >= pat.size() ) break;
                if ( ! isSpaceByte(ch) ) quality->push_back((char)(ch - qualityOffset));
        }
        in.p = p;
        if ( quality->size() < pat.size() )
                return false;
        {
                // findNextMarker
                void const * e = p < N ? memchr(B + p, marker, N - p) : 0;
                if ( e ) { in.p = static_cast<char const *>(e) - B + 1; foundnextmarker = true; }
                else in.p = N;
        }
        return true;
}

static bool findFirstMarker(Cursor & in, char marker)
{
        int c;
        while ( (c = in.get()) >= 0 && c != marker ) {}
        return c == marker;
}

static int detectQualityOffsetBuffer(FileBytes const & buf)
{
        Cursor in(buf);
        bool found = findFirstMarker(in, '@');
        std::string id, pat, q;
        while ( nextPattern(in, true, 0, found, id, pat, &q) )
                for ( size_t i = 0; i < q.size(); ++i )
                {
                        if ( q[i] <= 54 ) return 33;            // Sanger
                        else if ( q[i] >= 94 ) return 64;       // Illumina
                }
        return 0;
}

int detectQualityOffset(std::string const & filename)
{
        FileBytes buf;
        slurp(filename, buf);
        return detectQualityOffsetBuffer(buf);
}

namespace
{
        static size_t const NO_MARKER = ~(size_t)0;

        // Runs the reader's state machine from the marker at `start` (NO_MARKER = nothing to read) and appends the records
        // whose marker lies in front of `limit`.  Returns the position of the marker the next record starts at, or
        // NO_MARKER when the reader's loop is over (end of file, or a record it rejects: the reference stops there too).
        size_t parseRange(FileBytes const & buf, size_t start, size_t limit, bool fastq, int qualityOffset, ReadSet & out)
        {
                if ( start == NO_MARKER )
                        return NO_MARKER;
                Cursor in(buf);
                in.p = start + 1;
                bool found = true;
                size_t at = start;
                std::string id, pat, q;
                while ( at < limit )
                {
                        if ( ! nextPattern(in, fastq, qualityOffset, found, id, pat, fastq ? &q : 0) )
                                return NO_MARKER;
                        size_t const m0 = out.mapped.size();
                        out.mapped.resize(m0 + pat.size());
                        for ( size_t i = 0; i < pat.size(); ++i ) out.mapped[m0 + i] = mapChar((unsigned char)pat[i]);
                        if ( fastq )
                                out.quality.insert(out.quality.end(), q.begin(), q.begin() + pat.size());
                        out.offsets.push_back(out.mapped.size());
                        out.ids.push_back(id);
                        if ( ! found )
                                return NO_MARKER;
                        at = in.p - 1;
                }
                return at;
        }

        // where a record probably starts at or behind `from`: a marker at the start of a line (FASTQ: with a '+' line two
        // lines on).  Only a guess -- readPatternsBuffer checks it against the state machine's own position.
        size_t guessRecordStart(FileBytes const & buf, size_t from, bool fastq)
        {
                char const marker = fastq ? '@' : '>';
                for ( size_t p = from; p < buf.size(); ++p )
                {
                        if ( buf[p] != marker || (p && buf[p-1] != '\n') )
                                continue;
                        if ( ! fastq )
                                return p;
                        size_t a = p;
                        while ( a < buf.size() && buf[a] != '\n' ) ++a;          // end of the id line
                        size_t b = a + 1;
                        while ( b < buf.size() && buf[b] != '\n' ) ++b;          // end of the sequence line
                        if ( b + 1 < buf.size() && buf[b+1] == '+' )
                                return p;
                }
                return NO_MARKER;
        }

}

// FastAReader / FastQReader::getNextPatternUnlocked over the whole file (FastAReader.hpp:107-138, FastQReader.hpp:130-180),
// with a team of host threads: the file is cut into byte ranges, every thread guesses where the first record of its
// range starts and runs the reader's state machine from there; a range is accepted only if its guess is exactly the
// position the state machine reached at the end of the range in front of it -- otherwise that range is parsed again
// from the right position.  The result is the serial reader's, whatever the file looks like.
void readPatternsBuffer(FileBytes const & buf, bool fastq, int qualityOffset, ReadSet & out, unsigned int threads)
{
        out.mapped.clear(); out.quality.clear(); out.ids.clear();
        out.offsets.assign(1, 0);
        size_t chunk = std::max<size_t>(size_t(1) << 22, (buf.size() + (threads ? threads : 1) - 1) / (threads ? threads : 1));
        if ( char const * e = getenv("REAL_PARSE_CHUNK") ) chunk = std::max<size_t>(1, strtoull(e, 0, 10));       // tests
        size_t const nchunks = std::max<size_t>(1, (buf.size() + chunk - 1) / chunk);
        size_t first = NO_MARKER;
        {
                Cursor in(buf);
                if ( findFirstMarker(in, fastq ? '@' : '>') ) first = in.p - 1;
        }
        if ( nchunks == 1 || threads <= 1 )
        {
                out.mapped.reserve(buf.size());
                if ( fastq ) out.quality.reserve(buf.size() / 2);
                parseRange(buf, first, buf.size(), fastq, qualityOffset, out);
                return;
        }
        std::chrono::steady_clock::time_point const tp0 = std::chrono::steady_clock::now();
        std::vector<ReadSet> part(nchunks);
        std::vector<size_t> guess(nchunks, NO_MARKER), reached(nchunks, NO_MARKER);
        std::atomic<size_t> next(0);
        auto work = [&]()
        {
                for ( size_t c = next++; c < nchunks; c = next++ )
                {
                        size_t const lo = c * chunk, hi = std::min(buf.size(), lo + chunk);
                        guess[c] = c ? guessRecordStart(buf, lo, fastq) : first;
                        part[c].offsets.assign(1, 0);
                        if ( guess[c] != NO_MARKER && guess[c] < hi )
                                reached[c] = parseRange(buf, guess[c], hi, fastq, qualityOffset, part[c]);
                        else
                                reached[c] = guess[c];          // no record starts in this range
                }
        };
        {
                std::vector<std::thread> team;
                for ( unsigned int t = 0; t < std::min<size_t>(threads, nchunks); ++t ) team.push_back(std::thread(work));
                for ( size_t t = 0; t < team.size(); ++t ) team[t].join();
        }
        std::chrono::steady_clock::time_point const tp1 = std::chrono::steady_clock::now();
        // hand-over check, range by range (a range whose guess was wrong is parsed again here), then the pieces are
        // copied into place by the team
        std::vector<char> use(nchunks, 0);
        size_t pos = first;                                     // the marker the serial reader's next record starts at
        for ( size_t c = 0; c < nchunks && pos != NO_MARKER; ++c )
        {
                size_t const hi = std::min(buf.size(), c * chunk + chunk);
                if ( pos >= hi )
                        continue;                               // the record in front runs across this whole range
                if ( guess[c] != pos )
                {
                        ReadSet again;
                        again.offsets.assign(1, 0);
                        reached[c] = parseRange(buf, pos, hi, fastq, qualityOffset, again);
                        part[c].mapped.swap(again.mapped); part[c].quality.swap(again.quality); part[c].offsets.swap(again.offsets); part[c].ids.swap(again.ids);
                }
                use[c] = 1;
                pos = reached[c];
        }
        std::vector<size_t> mbase(nchunks + 1, 0), rbase(nchunks + 1, 0);
        for ( size_t c = 0; c < nchunks; ++c )
        {
                mbase[c+1] = mbase[c] + (use[c] ? part[c].mapped.size() : 0);
                rbase[c+1] = rbase[c] + (use[c] ? part[c].ids.size() : 0);
        }
        out.mapped.resize(mbase[nchunks]);
        if ( fastq ) out.quality.resize(mbase[nchunks]);
        out.offsets.resize(rbase[nchunks] + 1);
        out.ids.resize(rbase[nchunks]);
        next = 0;
        auto place = [&]()
        {
                for ( size_t c = next++; c < nchunks; c = next++ )
                {
                        if ( ! use[c] ) continue;
                        ReadSet & P = part[c];
                        if ( ! P.mapped.empty() ) memcpy(&out.mapped[mbase[c]], &P.mapped[0], P.mapped.size());
                        if ( fastq && ! P.quality.empty() ) memcpy(&out.quality[mbase[c]], &P.quality[0], P.quality.size());
                        for ( size_t i = 0; i < P.ids.size(); ++i )
                        {
                                out.offsets[rbase[c] + i + 1] = mbase[c] + P.offsets[i+1];
                                out.ids[rbase[c] + i].swap(P.ids[i]);
                        }
                        ReadSet().mapped.swap(P.mapped); ReadSet().quality.swap(P.quality);
                }
        };
        {
                std::vector<std::thread> team;
                for ( unsigned int t = 0; t < std::min<size_t>(threads, nchunks); ++t ) team.push_back(std::thread(place));
                for ( size_t t = 0; t < team.size(); ++t ) team[t].join();
        }
        if ( getenv("REAL_TIMING") )
                std::cerr << "[timing]   parse (" << nchunks << " ranges, " << threads << " threads) "
                          << std::chrono::duration<double>(tp1 - tp0).count() << " s, hand-over check + merge "
                          << std::chrono::duration<double>(std::chrono::steady_clock::now() - tp1).count() << " s" << std::endl;
}

void readPatterns(std::string const & filename, bool fastq, int qualityOffset, ReadSet & out, unsigned int threads)
{
        FileBytes buf;
        slurp(filename, buf);
        if ( ! threads ) threads = std::max(1u, std::thread::hardware_concurrency());
        readPatternsBuffer(buf, fastq, qualityOffset, out, threads);
}

void reorderLikeRewrite(ReadSet & reads)
{
        uint64_t const n = reads.size();
        std::vector< std::pair< std::pair<uint64_t,int>, uint64_t> > key(n);
        for ( uint64_t r = 0; r < n; ++r )
        {
                uint64_t const L = reads.offsets[r+1] - reads.offsets[r];
                int hasn = 0;
                for ( uint64_t i = reads.offsets[r]; i < reads.offsets[r+1]; ++i ) if ( reads.mapped[i] > 3 ) { hasn = 1; break; }
                key[r] = std::make_pair(std::make_pair(L, hasn), r);
        }
        std::sort(key.begin(), key.end());     // r is part of the key: file order inside a group
        ReadSet o;
        o.mapped.reserve(reads.mapped.size());
        if ( reads.quality.size() ) o.quality.reserve(reads.quality.size());
        o.offsets.assign(1, 0);
        for ( uint64_t i = 0; i < n; ++i )
        {
                uint64_t const r = key[i].second;
                o.mapped.insert(o.mapped.end(), reads.mapped.begin() + reads.offsets[r], reads.mapped.begin() + reads.offsets[r+1]);
                if ( reads.quality.size() )
                        o.quality.insert(o.quality.end(), reads.quality.begin() + reads.offsets[r], reads.quality.begin() + reads.offsets[r+1]);
                o.offsets.push_back(o.mapped.size());
                o.ids.push_back(reads.ids[r]);
        }
        reads.mapped.swap(o.mapped); reads.quality.swap(o.quality); reads.offsets.swap(o.offsets); reads.ids.swap(o.ids);
}


// ---- the reference's rewritten pattern file ---------------------------------------------------------

namespace
{
        inline void putU32(std::vector<char> & o, uint32_t v) { o.push_back((char)(v >> 24)); o.push_back((char)(v >> 16)); o.push_back((char)(v >> 8)); o.push_back((char)v); }
        inline void putU64(std::vector<char> & o, uint64_t v) { putU32(o, (uint32_t)(v >> 32)); putU32(o, (uint32_t)v); }
        // section header: byte count (magic included; TemporaryFile::writeContent writes the length of its file), magic
        inline void putSection(std::vector<char> & o, uint64_t payload, uint32_t magic)
        {
                uint64_t const bl = payload + 4;
                if ( bl >= 0xFFFFFFFFULL ) { putU32(o, 0xFFFFFFFFu); putU64(o, bl); }
                else putU32(o, (uint32_t)bl);
                putU32(o, magic);
        }
}

void writeRewritten(ReadSet const & reads, bool fastq, std::vector<char> & out)
{
        out.clear();
        uint64_t const n = reads.size();
        uint64_t r = 0;
        while ( r < n )
        {
                uint64_t const L = reads.offsets[r+1] - reads.offsets[r];
                // the reads of this length: wildcard free ones first [r, rn), then the others [rn, re)
                uint64_t rn = r, re = r;
                bool inN = false;
                while ( re < n && reads.offsets[re+1] - reads.offsets[re] == L )
                {
                        bool hasn = false;
                        for ( uint64_t i = reads.offsets[re]; i < reads.offsets[re+1]; ++i ) if ( reads.mapped[i] > 3 ) { hasn = true; break; }
                        if ( hasn ) inN = true;
                        else if ( inN ) throw std::runtime_error("writeRewritten: the reads are not in rewritten order");
                        if ( ! inN ) rn = re + 1;
                        ++re;
                }
                putU32(out, (uint32_t)L);
                uint64_t const q = fastq ? L : 0;
                for ( int part = 0; part < 2; ++part )
                {
                        uint64_t const a = part ? rn : r, b = part ? re : rn;
                        uint64_t const code = part ? (L + 1) / 2 : (L + 3) / 4;
                        putSection(out, (b - a) * (code + q), part ? 2u : 0u);
                        for ( uint64_t x = a; x < b; ++x )
                        {
                                const uint8_t * m = &reads.mapped[0] + reads.offsets[x];
                                if ( ! part )
                                        for ( uint64_t i = 0; i < L; i += 4 )
                                        {
                                                unsigned int c = 0;
                                                for ( uint64_t j = 0; j < 4 && i + j < L; ++j ) c |= (unsigned int)(m[i+j] & 3) << (6 - 2*j);
                                                out.push_back((char)c);
                                        }
                                else
                                        for ( uint64_t i = 0; i < L; i += 2 )
                                                out.push_back((char)((std::min<unsigned int>(m[i], 4) << 4) | (i + 1 < L ? std::min<unsigned int>(m[i+1], 4) : 0u)));
                                if ( fastq ) out.insert(out.end(), reads.quality.begin() + reads.offsets[x], reads.quality.begin() + reads.offsets[x+1]);
                        }
                        uint64_t idbytes = 0;
                        for ( uint64_t x = a; x < b; ++x ) idbytes += 2 + reads.ids[x].size();
                        putSection(out, idbytes, part ? 3u : 1u);
                        for ( uint64_t x = a; x < b; ++x )
                        {
                                out.push_back((char)(reads.ids[x].size() >> 8)); out.push_back((char)reads.ids[x].size());      // writeNumber2
                                out.insert(out.end(), reads.ids[x].begin(), reads.ids[x].end());
                        }
                }
                r = re;
        }
}

bool looksRewritten(FileBytes const & buf) { return buf.size() >= 8 && buf[0] == 0; }

// the sections of a rewritten file: per pattern length the wildcard-free part (0) and the part with wildcards (1)
namespace
{
        struct RewrittenPart { uint64_t L, data, ids, nreads; int part; };

        void rewrittenLayout(FileBytes const & buf, std::vector<RewrittenPart> & parts, bool & fastq)
        {
                parts.clear();
                uint64_t at = 0;
                uint64_t const size = buf.size();
                auto need = [&](uint64_t k) { if ( at + k > size ) throw std::runtime_error("rewritten pattern file: truncated"); };
                auto getU32 = [&]() -> uint32_t { need(4); const unsigned char * p = (const unsigned char *)&buf[at]; at += 4; return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; };
                auto getSection = [&](uint32_t magic) -> uint64_t
                {
                        uint64_t bl = getU32();
                        if ( bl == 0xFFFFFFFFULL ) { uint64_t const hi = getU32(), lo = getU32(); bl = (hi << 32) | lo; }
                        if ( bl < 4 || getU32() != magic ) throw std::runtime_error("rewritten pattern file: unexpected section");
                        need(bl - 4);
                        return bl - 4;
                };
                bool known = false;
                fastq = false;
                while ( at < size )
                {
                        uint64_t const L = getU32();
                        for ( int part = 0; part < 2; ++part )
                        {
                                RewrittenPart P;
                                P.L = L; P.part = part;
                                uint64_t const ndata = getSection(part ? 2u : 0u); P.data = at; at += ndata;
                                uint64_t const nids = getSection(part ? 3u : 1u); P.ids = at; at += nids;
                                // the ids tell how many reads the section holds (the record size tells FASTA from FASTQ)
                                uint64_t c = 0, x = P.ids;
                                uint64_t const end = P.ids + nids;
                                while ( x < end )
                                {
                                        if ( x + 2 > end ) throw std::runtime_error("rewritten pattern file: broken id section");
                                        x += 2 + (((uint64_t)(unsigned char)buf[x] << 8) | (unsigned char)buf[x+1]);
                                        ++c;
                                }
                                if ( x != end ) throw std::runtime_error("rewritten pattern file: broken id section");
                                P.nreads = c;
                                uint64_t const code = part ? (L + 1) / 2 : (L + 3) / 4;
                                if ( c && L )
                                {
                                        bool fq;
                                        if ( ndata == c * code ) fq = false;
                                        else if ( ndata == c * (code + L) ) fq = true;
                                        else throw std::runtime_error("rewritten pattern file: section size does not match its ids");
                                        if ( known && fq != fastq ) throw std::runtime_error("rewritten pattern file: mixed record kinds");
                                        known = true; fastq = fq;
                                }
                                else if ( ! c && ndata ) throw std::runtime_error("rewritten pattern file: section size does not match its ids");
                                parts.push_back(P);
                        }
                }
        }
}

bool rewrittenIsFastq(FileBytes const & buf)
{
        std::vector<RewrittenPart> parts; bool fastq = false;
        rewrittenLayout(buf, parts, fastq);
        return fastq;
}

void readRewritten(FileBytes const & buf, ReadSet & reads, bool & fastq)
{
        reads.mapped.clear(); reads.quality.clear(); reads.ids.clear(); reads.offsets.assign(1, 0);
        std::vector<RewrittenPart> parts;
        rewrittenLayout(buf, parts, fastq);
        for ( size_t g = 0; g < parts.size(); ++g )
        {
                RewrittenPart const & P = parts[g];
                uint64_t const L = P.L, code = P.part ? (L + 1) / 2 : (L + 3) / 4, rec = code + (fastq ? L : 0);
                uint64_t x = P.ids;
                for ( uint64_t i = 0; i < P.nreads; ++i )
                {
                        const unsigned char * d = (const unsigned char *)&buf[0] + P.data + i * rec;
                        size_t const o = reads.mapped.size();
                        reads.mapped.resize(o + L);
                        if ( ! P.part )
                                for ( uint64_t j = 0; j < L; ++j ) reads.mapped[o + j] = (d[j >> 2] >> (6 - 2 * (j & 3))) & 3;
                        else
                                for ( uint64_t j = 0; j < L; ++j ) reads.mapped[o + j] = std::min<unsigned int>((d[j >> 1] >> ((j & 1) ? 0 : 4)) & 15, 4);
                        if ( fastq ) reads.quality.insert(reads.quality.end(), d + code, d + code + L);
                        reads.offsets.push_back(reads.mapped.size());
                        uint64_t const il = ((uint64_t)(unsigned char)buf[x] << 8) | (unsigned char)buf[x+1];
                        reads.ids.push_back(std::string(&buf[0] + x + 2, &buf[0] + x + 2 + il));
                        x += 2 + il;
                }
        }
}

// The same file straight into what the device takes: the ACGT sections ARE the 2 bit/base layout of real_gpu_set_reads_packed
// (every read on a byte boundary, 4 bases per byte) and are copied as they stand -- one memcpy per section for FASTA, one per
// read for FASTQ, whose records carry their qualities behind the bases; the reads of the ACGTN sections are flagged and stored
// as A.  No byte-per-base copy of the reads is made.  ids: one byte string + offsets (real_gpu_set_read_ids).
void readRewrittenPacked(FileBytes const & buf, PackedReads & out, std::vector<uint8_t> & quality, std::vector<char> & idbytes, std::vector<uint64_t> & idoff, bool & fastq)
{
        std::vector<RewrittenPart> parts;
        rewrittenLayout(buf, parts, fastq);
        uint64_t n = 0, nbytes = 0, nbases = 0, nid = 0;
        for ( size_t g = 0; g < parts.size(); ++g )
        {
                n += parts[g].nreads; nbytes += parts[g].nreads * ((parts[g].L + 3) / 4); nbases += parts[g].nreads * parts[g].L;
        }
        out.packed.assign(nbytes + 8, 0);
        out.byte_offsets.assign(n + 1, 0); out.lengths.assign(n, 0); out.wildcard.assign(n, 0);
        quality.clear();
        if ( fastq ) quality.resize(nbases);
        idoff.assign(n + 1, 0);
        uint64_t r = 0, bo = 0, qo = 0;
        for ( size_t g = 0; g < parts.size(); ++g )
        {
                RewrittenPart const & P = parts[g];
                uint64_t const L = P.L, pk = (L + 3) / 4, code = P.part ? (L + 1) / 2 : pk, rec = code + (fastq ? L : 0);
                const unsigned char * d = (const unsigned char *)&buf[0] + P.data;
                if ( ! P.part && ! fastq && P.nreads )
                        memcpy(&out.packed[bo], d, P.nreads * pk);                 // the section as it stands
                uint64_t x = P.ids;
                for ( uint64_t i = 0; i < P.nreads; ++i, ++r )
                {
                        if ( ! P.part && fastq ) memcpy(&out.packed[bo], d + i * rec, pk);
                        if ( fastq ) { memcpy(&quality[qo], d + i * rec + code, L); qo += L; }
                        out.lengths[r] = (uint32_t)L; out.wildcard[r] = P.part ? 1 : 0;
                        bo += pk; out.byte_offsets[r+1] = bo;
                        uint64_t const il = ((uint64_t)(unsigned char)buf[x] << 8) | (unsigned char)buf[x+1];
                        idoff[r+1] = idoff[r] + il; nid += il;
                        x += 2 + il;
                }
        }
        idbytes.resize(nid + 1);
        r = 0;
        for ( size_t g = 0; g < parts.size(); ++g )
        {
                uint64_t x = parts[g].ids;
                for ( uint64_t i = 0; i < parts[g].nreads; ++i, ++r )
                {
                        uint64_t const il = idoff[r+1] - idoff[r];
                        memcpy(&idbytes[idoff[r]], &buf[0] + x + 2, il);
                        x += 2 + il;
                }
        }
}

// ---------------------------------------------------------------------------------------------
// drivers
// ---------------------------------------------------------------------------------------------

void packReads(ReadSet const & reads, PackedReads & out, unsigned int threads)
{
        uint64_t const n = reads.size();
        out.byte_offsets.assign(n + 1, 0);
        out.lengths.assign(n, 0);
        out.wildcard.assign(n, 0);
        for ( uint64_t r = 0; r < n; ++r )
        {
                uint64_t const L = reads.offsets[r+1] - reads.offsets[r];
                out.lengths[r] = (uint32_t)L;
                out.byte_offsets[r+1] = out.byte_offsets[r] + (L + 3) / 4;
        }
        out.packed.assign(out.byte_offsets[n] + 8, 0);
        if ( threads < 1 ) threads = 1;
        threads = (unsigned int)std::min<uint64_t>(threads, std::max<uint64_t>(1, n / 4096));
        auto work = [&reads, &out, n, threads](unsigned int t)
        {
                uint64_t const a = n * t / threads, b = n * (t + 1) / threads;
                for ( uint64_t r = a; r < b; ++r )
                {
                        uint8_t const * m = &reads.mapped[0] + reads.offsets[r];
                        uint32_t const L = out.lengths[r];
                        uint8_t * q = &out.packed[0] + out.byte_offsets[r];
                        uint8_t bad = 0;
                        uint32_t i = 0;
                        for ( ; i + 4 <= L; i += 4 )
                        {
                                bad |= (m[i] | m[i+1] | m[i+2] | m[i+3]) & 0xFC;
                                *q++ = (uint8_t)(((m[i] & 3) << 6) | ((m[i+1] & 3) << 4) | ((m[i+2] & 3) << 2) | (m[i+3] & 3));
                        }
                        if ( i < L )
                        {
                                uint8_t v = 0;
                                for ( uint32_t j = 0; i + j < L; ++j ) { bad |= m[i+j] & 0xFC; v |= (uint8_t)((m[i+j] & 3) << (6 - 2 * j)); }
                                *q++ = v;
                        }
                        if ( bad )
                        {
                                // a wildcard: flagged, its bases stored as A (the flag is what keeps the read from matching)
                                out.wildcard[r] = 1;
                                memset(&out.packed[0] + out.byte_offsets[r], 0, (L + 3) / 4);
                        }
                }
        };
        if ( threads == 1 ) { work(0); return; }
        std::vector<std::thread> team;
        for ( unsigned int t = 0; t < threads; ++t ) team.push_back(std::thread(work, t));
        for ( size_t t = 0; t < team.size(); ++t ) team[t].join();
}

namespace
{
        struct Gpu
        {
                real_gpu * h;
                Gpu() : h(0) {}
                ~Gpu() { if ( h ) real_gpu_destroy(h); }
                void check(int rc, char const * what) const
                {
                        if ( rc != REAL_GPU_OK )
                                throw std::runtime_error(std::string(what) + ": " + (h ? real_gpu_last_error(h) : "library error"));
                }
        };

        struct Output
        {
                FILE * f; bool own;
                explicit Output(std::string const & name) : f(0), own(false)
                {
                        if ( name == "-" ) f = stdout;
                        else { f = fopen(name.c_str(), "wb"); own = true; }
                        if ( ! f ) throw std::runtime_error("Failed to open output file " + name);
                }
                ~Output() { if ( f ) { fflush(f); if ( own ) fclose(f); } }
                void write(std::string const & s) { if ( fwrite(s.data(), 1, s.size(), f) != s.size() ) throw std::runtime_error("write failed"); }
        };

        inline void appendUnsigned(std::string & o, uint64_t v)
        {
                char tmp[24]; int n = 0;
                do { tmp[n++] = (char)('0' + v % 10); v /= 10; } while ( v );
                while ( n ) o.push_back(tmp[--n]);
        }

        // one result line (matchAllImplementation.cpp:485-510, matchUniqueImplementation.cpp:267-288); the reference prints
        // through an ostream with default flags: integers in decimal, the float score like printf's %g
        void formatLine(std::string & o, ReadSet const & reads, uint64_t r, bool inverted, bool scores, float score,
                        std::string const & recname, uint64_t pos_in_record, unsigned int k)
        {
                static char const remap[5] = { 'A', 'C', 'G', 'T', 'N' };
                uint64_t const b = reads.offsets[r], e = reads.offsets[r+1];
                o += reads.ids[r]; o.push_back('\t');
                size_t const at = o.size();
                o.resize(at + (e - b));
                if ( ! inverted )
                        for ( uint64_t i = b; i < e; ++i ) o[at + (i - b)] = remap[std::min<int>(reads.mapped[i], 4)];
                else
                        for ( uint64_t i = e; i > b; --i ) { int const c = reads.mapped[i-1]; o[at + (e - i)] = remap[c < 4 ? 3 - c : 4]; }
                o.push_back('\t');
                if ( scores )
                {
                        char tmp[32];
                        int const n = snprintf(tmp, sizeof(tmp), "%g", (double)score);
                        o.append(tmp, n);
                }
                o += "\t1\ta\t";
                appendUnsigned(o, e - b);
                o += inverted ? "\t-\t" : "\t+\t";
                o += recname; o.push_back('\t');
                appendUnsigned(o, pos_in_record);
                o += "\t\t";
                appendUnsigned(o, k);
                o.push_back('\n');
        }

        // Formats the items [0,n) with `threads` host threads and writes the pieces in item order, wave by wave (the
        // reference formats and prints serially: matchUniqueImplementation.cpp:1455-1486).  fmt(o, i) appends the lines
        // of item i to o and returns how many it wrote.  The bytes written are those of the serial loop.
        template<typename F>
        uint64_t formatParallel(uint64_t n, unsigned int threads, Output & out, F fmt)
        {
                if ( threads < 1 ) threads = 1;
                uint64_t per = 1u << 16;                         // items per thread and wave (REAL_FORMAT_CHUNK: tests)
                if ( char const * e = getenv("REAL_FORMAT_CHUNK") ) per = std::max<uint64_t>(1, strtoull(e, 0, 10));
                uint64_t total = 0;
                std::vector<std::string> piece(threads);
                std::vector<uint64_t> lines(threads);
                for ( uint64_t w0 = 0; w0 < n; w0 += per * threads )
                {
                        uint64_t const w1 = std::min<uint64_t>(n, w0 + per * threads);
                        unsigned int const used = (unsigned int)((w1 - w0 + per - 1) / per);
                        std::vector<std::thread> team;
                        for ( unsigned int t = 0; t < used; ++t )
                        {
                                uint64_t const a = w0 + t * per, b = std::min<uint64_t>(w1, a + per);
                                auto work = [&piece, &lines, &fmt, t, a, b]()
                                {
                                        std::string o;
                                        o.reserve((b - a) * 192);
                                        uint64_t c = 0;
                                        for ( uint64_t i = a; i < b; ++i ) c += fmt(o, i);
                                        piece[t].swap(o);
                                        lines[t] = c;
                                };
                                if ( used == 1 ) work(); else team.push_back(std::thread(work));
                        }
                        for ( size_t t = 0; t < team.size(); ++t ) team[t].join();
                        for ( unsigned int t = 0; t < used; ++t ) { out.write(piece[t]); total += lines[t]; }
                }
                return total;
        }

        unsigned int hostThreads(RealOptions const & opts)
        {
                if ( opts.threads > 0 ) return (unsigned int)opts.threads;
                unsigned int const hc = std::thread::hardware_concurrency();
                return hc ? hc : 1;
        }

        struct PhaseTimer
        {
                std::chrono::steady_clock::time_point t0;
                PhaseTimer() : t0(std::chrono::steady_clock::now()) {}
                void lap(char const * what)
                {
                        std::chrono::steady_clock::time_point const t1 = std::chrono::steady_clock::now();
                        if ( getenv("REAL_TIMING") )
                                std::cerr << "[timing] " << what << " " << std::chrono::duration<double>(t1 - t0).count() << " s" << std::endl;
                        t0 = t1;
                }
        };

        void createHandle(Gpu & G, RealOptions const & opts, std::vector<double> & ll, int device = -1)
        {
                real_gpu_params P;
                memset(&P, 0, sizeof(P));
                P.struct_size = sizeof(P);
                P.device = device >= 0 ? device : opts.device;
                P.seedl = opts.seedl; P.seedkmax = opts.seedkmax; P.totalkmax = opts.totalkmax; P.scores = opts.scores ? 1 : 0;
                P.filter_mult = opts.filter_mult;
                if ( opts.scores || opts.gaps )
                {
                        ll.resize(1024);
                        buildScoringTable(opts.similarity, opts.gc, opts.trans, opts.err, opts.gcmut_bias, &ll[0]);
                        P.ll_table = &ll[0];
                }
                int const rc = real_gpu_create(&P, &G.h);
                if ( rc != REAL_GPU_OK )
                        throw std::runtime_error("real_gpu_create failed (no CUDA device or unsupported option); there is no CPU fallback");
        }

        // getText (getText.hpp:31-58) for the device: the bytes of the file go to real_gpu_set_text_fasta, which parses and
        // packs them there and sets the text; the record table comes back for the output lines.  T.words / T.nmask stay empty.
        // REAL_TEXT_LOADER=host selects the host parser + real_gpu_set_text instead (timing comparisons, tests).
        // Returns false when the file holds nothing to match against (too short for the seed, or over the library's limits
        // when skip_over_limits is set).
        bool setText(Gpu & G, RealOptions const & opts, uint32_t fileid, std::string const & filename, TextFile & T, bool skip_over_limits)
        {
                char const * const loader = getenv("REAL_TEXT_LOADER");
                if ( loader && std::string(loader) == "host" )
                {
                        getText(filename, T);
                        if ( skip_over_limits && (fileid >= 64 || T.n >= (1ULL << 35)) )
                                return false;
                        if ( T.n < (uint64_t)opts.seedl )
                                return false;
                        std::vector<uint64_t> const starts = T.starts();
                        G.check(real_gpu_set_text(G.h, fileid, &T.words[0], &T.nmask[0], T.n, 0, T.n, 0, T.n, &starts[0], (uint32_t)(starts.size() - 1)), "set_text");
                        return true;
                }
                std::cerr << "Computing length of file " << filename << "...";
                FileBytes buf;
                slurp(filename, buf);
                T.ranges.clear(); T.words.clear(); T.nmask.clear(); T.n = 0;
                if ( skip_over_limits && fileid >= 64 )
                        return false;
                uint64_t n = 0, nrec = 0;
                int const rc = real_gpu_set_text_fasta(G.h, fileid, buf.empty() ? 0 : &buf[0], buf.size(), &n, &nrec);
                T.n = n;
                std::cerr << "done, length is " << n << std::endl;
                if ( rc == REAL_GPU_E_LIMIT && skip_over_limits )
                {
                        T.ranges.push_back(std::pair<std::string, uint64_t>("terminal", n));
                        return false;
                }
                G.check(rc, "set_text_fasta");
                if ( n && nrec )
                {
                        std::vector<uint64_t> starts(nrec + 1), ends(nrec);
                        G.check(real_gpu_get_text_records(G.h, &starts[0], &ends[0]), "get_text_records");
                        T.ranges.reserve(nrec + 1);
                        for ( uint64_t r = 0; r < nrec; ++r )
                        {
                                uint64_t b = ends[r];
                                while ( b > 0 && buf[b-1] != '>' ) --b;         // the name starts behind the line's last '>' (countReads.cpp:47-51)
                                T.ranges.push_back(std::pair<std::string, uint64_t>(std::string(&buf[0] + b, &buf[0] + ends[r]), starts[r]));
                        }
                }
                T.ranges.push_back(std::pair<std::string, uint64_t>("terminal", n));
                if ( n < (uint64_t)opts.seedl )
                        return false;
                if ( ! nrec )
                        throw std::runtime_error("set_text: null pointer or no records");
                return true;
        }

        // runs fn(i) for i in [0,n) on n host threads (one per GPU handle) and rethrows the first error
        template<typename F>
        void parallelFor(unsigned int n, F fn)
        {
                if ( n == 1 ) { fn(0u); return; }
                std::vector<std::thread> team;
                std::vector<std::exception_ptr> err(n);
                for ( unsigned int i = 0; i < n; ++i )
                        team.push_back(std::thread([&fn, &err, i]() { try { fn(i); } catch ( ... ) { err[i] = std::current_exception(); } }));
                for ( unsigned int i = 0; i < n; ++i ) team[i].join();
                for ( unsigned int i = 0; i < n; ++i ) if ( err[i] ) std::rethrow_exception(err[i]);
        }

        // The GPUs of the box in one process (real.cpp:203-230 is one process): REAL_GPUS handles, handle i on device
        // (REAL_GPU_DEVICE + i) modulo the devices present, each indexing and probing 1/N of the signature space
        // (real_gpu_set_bucket_shard) against the whole text.  matchAll: the rows of the handles are merged on the host;
        // matchUnique: the per-read states are folded over peer memory (real_gpu_fold_unique_group).  One host thread per handle
        // drives the uploads and the scans.  The order dependent folds (scores, gapped pass) need all hits of a read in one
        // place and stay on one handle.
        struct GpuTeam
        {
                std::vector<Gpu> g;
                std::vector<double> ll;
                std::thread starter; std::exception_ptr starter_err;
                unsigned int size() const { return (unsigned int)g.size(); }
                GpuTeam(RealOptions const & opts, bool order_dependent)
                {
                        unsigned int n = (unsigned int)opts.ngpus;
                        if ( n > 8 ) n = 8;
                        if ( order_dependent && n > 1 )
                        {
                                std::cerr << "REAL_GPUS=" << n << " ignored: the fold of this mode depends on the visiting order and runs on one GPU" << std::endl;
                                n = 1;
                        }
                        g.resize(n);
                        // real_gpu_create on threads of their own (context creation and module load take a few hundred
                        // milliseconds) while the pattern file is read
                        starter = std::thread([this, &opts]()
                        {
                                try
                                {
                                        int const ndev = std::max(1, real_gpu_device_count());
                                        std::vector<double> * llp = &ll;
                                        std::vector< std::vector<double> > lls(g.size());
                                        parallelFor((unsigned int)g.size(), [this, &opts, &lls, ndev](unsigned int i)
                                        { createHandle(g[i], opts, lls[i], (opts.device + (int)i) % ndev); });
                                        (void)llp;
                                }
                                catch ( ... ) { starter_err = std::current_exception(); }
                        });
                }
                void wait()
                {
                        if ( starter.joinable() ) starter.join();
                        if ( starter_err ) { std::exception_ptr e = starter_err; starter_err = nullptr; std::rethrow_exception(e); }
                }
                ~GpuTeam() { if ( starter.joinable() ) starter.join(); }
                // after wait(): bucket shards (before the reads are set) ...
                void shard()
                {
                        unsigned int const n = size();
                        if ( n == 1 ) return;
                        for ( unsigned int i = 0; i < n; ++i ) g[i].check(real_gpu_set_bucket_shard(g[i].h, i, n), "set_bucket_shard");
                }
                // ... and, for matchUnique, the fold windows (once the number of reads is known)
                void connectFold(uint64_t nreads)
                {
                        unsigned int const n = size();
                        if ( n == 1 ) return;
                        for ( unsigned int i = 0; i < n; ++i ) g[i].check(real_gpu_fold_init(g[i].h, i, n, nreads, 0), "fold_init");
                        std::vector<real_gpu *> hs(n);
                        for ( unsigned int i = 0; i < n; ++i ) hs[i] = g[i].h;
                        for ( unsigned int i = 0; i < n; ++i ) g[i].check(real_gpu_fold_connect_local(g[i].h, &hs[0]), "fold_connect_local");
                }
                void connect(uint64_t nreads, bool unique)
                {
                        shard();
                        if ( unique ) connectFold(nreads);
                }
                // The pattern file parsed on the device (K0 in pattern-file mode, real_gpu_set_reads_fasta): every handle is given the
                // bytes of the FASTA file and parses, packs and orders the reads itself; the ids stay on the device for the formatter.
                uint64_t setReadsFasta(FileBytes const & buf, bool rewrite_order)
                {
                        std::vector<uint64_t> n(size(), 0);
                        parallelFor(size(), [this, &buf, &n, rewrite_order](unsigned int i)
                        { g[i].check(real_gpu_set_reads_fasta(g[i].h, buf.empty() ? 0 : &buf[0], buf.size(), rewrite_order ? 1u : 0u, &n[i]), "set_reads_fasta"); });
                        return n[0];
                }
                void setReadsPacked(PackedReads const & P, std::vector<uint8_t> const & quality)
                {
                        uint64_t const n = P.lengths.size();
                        parallelFor(size(), [this, &P, &quality, n](unsigned int i)
                        {
                                g[i].check(real_gpu_set_reads_packed(g[i].h, &P.packed[0], &P.byte_offsets[0], P.lengths.empty() ? 0 : &P.lengths[0], 0,
                                                                     P.wildcard.empty() ? 0 : &P.wildcard[0], quality.empty() ? 0 : &quality[0], n), "set_reads_packed");
                        });
                }
                void setReads(ReadSet const & reads, PackedReads const & P)
                {
                        parallelFor(size(), [this, &reads, &P](unsigned int i)
                        {
                                g[i].check(real_gpu_set_reads_packed(g[i].h, &P.packed[0], &P.byte_offsets[0], P.lengths.empty() ? 0 : &P.lengths[0], 0,
                                                                     P.wildcard.empty() ? 0 : &P.wildcard[0], reads.quality.empty() ? 0 : &reads.quality[0], reads.size()), "set_reads_packed");
                        });
                }
                void foldUnique()
                {
                        if ( size() == 1 ) return;
                        std::vector<real_gpu *> hs(size());
                        for ( unsigned int i = 0; i < size(); ++i ) hs[i] = g[i].h;
                        int const rc = real_gpu_fold_unique_group(&hs[0], size());
                        if ( rc != REAL_GPU_OK )
                                for ( unsigned int i = 0; i < size(); ++i ) g[i].check(real_gpu_last_error(g[i].h)[0] ? rc : REAL_GPU_OK, "fold_unique_group");
                        g[0].check(rc, "fold_unique_group");
                }
        };

        // setText for every handle of the team (the bytes of the file are read once; every handle parses them on its GPU)
        bool setTextTeam(GpuTeam & team, RealOptions const & opts, uint32_t fileid, std::string const & filename, TextFile & T, bool skip_over_limits)
        {
                if ( team.size() == 1 )
                        return setText(team.g[0], opts, fileid, filename, T, skip_over_limits);
                bool const usable = setText(team.g[0], opts, fileid, filename, T, skip_over_limits);
                if ( ! usable )
                        return false;
                // the other handles: the same call without the messages (its stderr lines are part of the stock behaviour, once)
                FileBytes buf;
                char const * const loader = getenv("REAL_TEXT_LOADER");
                bool const host_loader = loader && std::string(loader) == "host";
                std::vector<uint64_t> const starts = host_loader ? T.starts() : std::vector<uint64_t>();
                if ( ! host_loader ) slurp(filename, buf);
                parallelFor(team.size() - 1, [&](unsigned int k)
                {
                        Gpu & G = team.g[k + 1];
                        if ( host_loader )
                                G.check(real_gpu_set_text(G.h, fileid, &T.words[0], &T.nmask[0], T.n, 0, T.n, 0, T.n, &starts[0], (uint32_t)(starts.size() - 1)), "set_text");
                        else
                        {
                                uint64_t n = 0, nrec = 0;
                                G.check(real_gpu_set_text_fasta(G.h, fileid, buf.empty() ? 0 : &buf[0], buf.size(), &n, &nrec), "set_text_fasta");
                        }
                });
                return true;
        }

        void loadReads(RealOptions const & opts, ReadSet & reads)
        {
                FileBytes buf;
                PhaseTimer PT;
                if ( opts.rewritten_input )
                {
                        // -p named a rewritten pattern file (the host formatter needs the reads one byte per base)
                        buf.open(opts.patternfilename);
                        bool fq = false;
                        readRewritten(buf, reads, fq);
                        std::cerr << "Number of patterns is " << reads.size() << std::endl;
                        return;
                }
                if ( opts.stdin_bytes )
                        buf.adopt(*opts.stdin_bytes);
                else
                slurp(opts.patternfilename, buf);
                PT.lap("  read pattern file");
                int qualityOffset = 0;
                if ( opts.fastq && opts.stdin_bytes && ! opts.qualityOffset )
                {
                        // real.cpp:248-257
                        std::cerr << "WARNING: automatic quality offset detection not supported when" << std::endl;
                        std::cerr << "         reading patterns from standard input. Assuming input" << std::endl;
                        std::cerr << "         was produced by an Illumina  GA (i.e. -Q 64)" << std::endl;
                        qualityOffset = 64;
                }
                else if ( opts.fastq )
                {
                        qualityOffset = opts.qualityOffset ? opts.qualityOffset : detectQualityOffsetBuffer(buf);
                        if ( ! qualityOffset )
                                throw std::runtime_error("Unable to automatically detect FastQ quality format.");
                }
                readPatternsBuffer(buf, opts.fastq, qualityOffset, reads, hostThreads(opts));
                std::cerr << "Number of patterns is " << reads.size() << std::endl;
        }

        // REAL_FORMAT=host: the lines are assembled by the host team (formatLine) instead of on the device (tests, timing)
        bool deviceFormat()
        {
                char const * const e = getenv("REAL_FORMAT");
                return ! (e && std::string(e) == "host");
        }

        // REAL_READS_LOADER=host: the pattern file is parsed by the host team instead of on the device.  The device reader takes
        // FASTA files (FASTQ needs the length-counted quality block and stays with the host team), and only when the lines are
        // formatted on the device too -- the host formatter needs the reads in host memory.
        bool deviceReadsLoader(RealOptions const & opts, bool devfmt)
        {
                char const * const e = getenv("REAL_READS_LOADER");
                if ( e && std::string(e) == "host" ) return false;
                return devfmt && ! opts.fastq && ! opts.rewritten_input && ! getenv("REAL_KEEP_REWRITTEN");
        }

        // the bytes of the pattern file (or of standard input)
        void patternBytes(RealOptions const & opts, FileBytes & buf)
        {
                if ( opts.stdin_bytes ) buf.adopt(*opts.stdin_bytes);
                else buf.open(opts.patternfilename);
        }

        // ids [lo, hi) out of one byte string + offsets
        void setReadIdsBlob(Gpu & G, std::vector<char> const & bytes, std::vector<uint64_t> const & off, uint64_t lo, uint64_t hi)
        {
                G.check(real_gpu_set_read_ids(G.h, lo, hi - lo, &bytes[0], &off[lo]), "set_read_ids");
        }

        // ids of the reads [lo, hi) as one byte string + offsets, for real_gpu_set_read_ids
        void setReadIds(Gpu & G, ReadSet const & reads, uint64_t lo, uint64_t hi)
        {
                std::vector<uint64_t> off(hi - lo + 1, 0);
                for ( uint64_t r = lo; r < hi; ++r ) off[r - lo + 1] = off[r - lo] + reads.ids[r].size();
                std::vector<char> bytes(off[hi - lo] + 1);
                for ( uint64_t r = lo; r < hi; ++r ) memcpy(&bytes[off[r - lo]], reads.ids[r].data(), reads.ids[r].size());
                G.check(real_gpu_set_read_ids(G.h, lo, hi - lo, &bytes[0], &off[0]), "set_read_ids");
        }

        // names and start offsets of the records of file fi (T.ranges without its terminal entry), for the lines that name them
        void setRecordNames(Gpu & G, uint32_t fi, std::vector< std::pair<std::string, uint64_t> > const & ranges)
        {
                if ( ranges.size() < 2 ) return;
                size_t const n = ranges.size() - 1;
                std::vector<uint64_t> off(n + 1, 0), starts(n);
                for ( size_t r = 0; r < n; ++r ) { off[r+1] = off[r] + ranges[r].first.size(); starts[r] = ranges[r].second; }
                std::vector<char> bytes(off[n] + 1);
                for ( size_t r = 0; r < n; ++r ) memcpy(&bytes[off[r]], ranges[r].first.data(), ranges[r].first.size());
                G.check(real_gpu_set_record_names(G.h, fi, (uint32_t)n, &bytes[0], &off[0], &starts[0]), "set_record_names");
        }

        // Writes the batches a formatter hands out, one behind the other, while the next batch is being formatted: next(&bytes,
        // &nbytes) formats the next batch into library-owned memory that stays valid until its next-but-one call, and returns
        // false when there is none left.
        // one batch to the output: a regular file takes it as four slices written side by side at their offsets (a single
        // writer copies at ~2.5 GB/s into the page cache: 3.1 of the 6.8 s of a run on C3-size inputs); a pipe or a device in order
        void writeBytes(Output & out, uint64_t at, char const * b, uint64_t n)
        {
                struct stat st;
                bool const regular = out.own && fstat(fileno(out.f), &st) == 0 && S_ISREG(st.st_mode);
                if ( ! regular )
                {
                        if ( fwrite(b, 1, n, out.f) != n ) throw std::runtime_error("write failed");
                        return;
                }
                int const fd = fileno(out.f);
                unsigned int const parts = n >= (32u << 20) ? 4u : 1u;
                std::atomic<int> failed(0);
                auto slice = [fd, at, b, n, parts, &failed](unsigned int p)
                {
                        uint64_t o = n * p / parts; uint64_t const e = n * (p + 1) / parts;
                        while ( o < e )
                        {
                                ssize_t const w = pwrite(fd, b + o, e - o, (off_t)(at + o));
                                if ( w <= 0 ) { failed = 1; return; }
                                o += (uint64_t)w;
                        }
                };
                std::vector<std::thread> team;
                for ( unsigned int p = 1; p < parts; ++p ) team.push_back(std::thread(slice, p));
                slice(0);
                for ( size_t t = 0; t < team.size(); ++t ) team[t].join();
                if ( failed ) throw std::runtime_error("write failed");
        }

        // Writes the batches a formatter hands out, one behind the other, while the next batch is being formatted: next(&bytes,
        // &nbytes) formats the next batch into library-owned memory that stays valid until its next-but-one call, and returns
        // false when there is none left.  `at` = bytes written to the output so far (kept by the caller across several calls).
        template<typename F>
        void writeBatches(Output & out, uint64_t & at, F next)
        {
                std::thread writer; std::exception_ptr err;
                char const * bytes = 0; uint64_t nbytes = 0;
                fflush(out.f);
                while ( next(&bytes, &nbytes) )
                {
                        if ( writer.joinable() ) writer.join();
                        if ( err ) std::rethrow_exception(err);
                        if ( ! nbytes ) continue;
                        char const * const b = bytes; uint64_t const n = nbytes, pos = at;
                        at += n;
                        writer = std::thread([&out, &err, b, n, pos]()
                        {
                                try { writeBytes(out, pos, b, n); }
                                catch ( ... ) { err = std::current_exception(); }
                        });
                }
                if ( writer.joinable() ) writer.join();
                if ( err ) std::rethrow_exception(err);
        }
}

// (patid, k, pos, file, frag, score, inverted): MatchPosAndError::operator< per read (matchAllImplementation.cpp:122-136)
static bool hitBefore(real_gpu_hit const & a, real_gpu_hit const & b)
{
        if ( a.patid != b.patid ) return a.patid < b.patid;
        if ( a.k != b.k ) return a.k < b.k;
        if ( a.pos != b.pos ) return a.pos < b.pos;
        if ( a.file != b.file ) return a.file < b.file;
        if ( a.frag != b.frag ) return a.frag < b.frag;
        if ( a.score != b.score ) return a.score < b.score;
        return a.inverted < b.inverted;
}

int doMatchingAll(RealOptions const & opts)
{
        PhaseTimer PT;
        GpuTeam team(opts, false);           // the CUDA contexts come up while the pattern file is read
        bool const devfmt = deviceFormat() && team.size() == 1;       // several handles: their rows are merged (and formatted) on the host
        ReadSet reads;
        std::vector<std::string> filenames;
        if ( deviceReadsLoader(opts, devfmt) )
        {
                // the pattern file goes to the device as it is: parsed, packed and indexed there, ids kept there for the formatter
                FileBytes buf;
                patternBytes(opts, buf);
                PT.lap("read patterns");
                getFileList(opts.textfilename, filenames, ".fa");
                team.wait();
                team.shard();
                uint64_t const n = team.setReadsFasta(buf, false);
                std::cerr << "Number of patterns is " << n << std::endl;
        }
        else if ( opts.rewritten_input && devfmt )
        {
                // a kept rewritten pattern file: its 2 bit/base sections go to the device as they stand
                FileBytes buf;
                buf.open(opts.patternfilename);
                PackedReads packed; std::vector<uint8_t> quality; std::vector<char> idbytes; std::vector<uint64_t> idoff; bool fq = false;
                readRewrittenPacked(buf, packed, quality, idbytes, idoff, fq);
                std::cerr << "Number of patterns is " << packed.lengths.size() << std::endl;
                PT.lap("read patterns");
                getFileList(opts.textfilename, filenames, ".fa");
                team.wait();
                team.connect(packed.lengths.size(), false);
                team.setReadsPacked(packed, quality);
                setReadIdsBlob(team.g[0], idbytes, idoff, 0, packed.lengths.size());
        }
        else
        {
                loadReads(opts, reads);
                PT.lap("read patterns");            // (the stock -u 0 path parses FASTQ files with the FASTA reader, real.cpp:325-328; here FASTQ is honoured)
                PackedReads packed;
                packReads(reads, packed, hostThreads(opts));
                getFileList(opts.textfilename, filenames, ".fa");
                team.wait();
                team.connect(reads.size(), false);
                team.setReads(reads, packed);
                if ( devfmt ) setReadIds(team.g[0], reads, 0, reads.size());
        }
        PT.lap("create + set_reads");
        Output out(opts.outputfilename);
        uint64_t written = 0;
        for ( size_t fi = 0; fi < filenames.size(); ++fi )
        {
                TextFile T;
                // (file slot 0 for every file: the rows of matchAll are printed file by file and their file field is not part of a
                // line, so the 64-file limit of the UniqueMatchInfo word does not apply here -- the reference has none either)
                if ( ! setTextTeam(team, opts, 0, filenames[fi], T, false) )
                {
                        std::cerr << "file " << filenames[fi] << " is too short for seed length " << opts.seedl << std::endl;
                        continue;
                }
                std::vector<real_gpu_hit const *> part(team.size(), (real_gpu_hit const *)0);
                std::vector<uint64_t> npart(team.size(), 0);
                parallelFor(team.size(), [&team, &part, &npart](unsigned int i)
                { team.g[i].check(real_gpu_match_all(team.g[i].h, &part[i], &npart[i]), "match_all"); });
                // one handle: its rows are in the order of the output already; several: every hit was found by exactly one of them,
                // the rows are merged into that order here
                real_gpu_hit const * hits = part[0]; uint64_t nhits = npart[0];
                std::vector<real_gpu_hit> merged;
                if ( team.size() > 1 )
                {
                        nhits = 0;
                        for ( unsigned int i = 0; i < team.size(); ++i ) nhits += npart[i];
                        merged.reserve(nhits);
                        for ( unsigned int i = 0; i < team.size(); ++i ) merged.insert(merged.end(), part[i], part[i] + npart[i]);
                        std::sort(merged.begin(), merged.end(), hitBefore);
                        hits = merged.empty() ? 0 : &merged[0];
                }
                PT.lap("text + match_all");
                bool const scores = opts.scores;
                if ( devfmt )
                {
                        // the lines are formatted on the device, batch by batch, and written while the next batch is formatted
                        Gpu & G = team.g[0];
                        setRecordNames(G, 0, T.ranges);
                        uint64_t const per = 1u << 18;
                        uint64_t at = 0;
                        writeBatches(out, written, [&G, &at, nhits, per](char const ** bytes, uint64_t * nbytes) -> bool
                        {
                                if ( at >= nhits ) return false;
                                uint64_t const c = std::min<uint64_t>(per, nhits - at);
                                G.check(real_gpu_format_all(G.h, at, c, bytes, nbytes), "format_all");
                                at += c;
                                return true;
                        });
                        PT.lap("format + write");
                        continue;
                }
                formatParallel(nhits, hostThreads(opts), out, [&reads, &T, hits, scores](std::string & o, uint64_t i) -> uint64_t
                {
                        real_gpu_hit const & H = hits[i];
                        formatLine(o, reads, H.patid, H.inverted != 0, scores, H.score, T.ranges[H.frag].first, H.pos - T.ranges[H.frag].second + 1, H.k);
                        return 1;
                });                         // (complete: the stock driver never flushes the tail of a block, matchAllImplementation.cpp:512-517)
                PT.lap("format + write");
        }
        return EXIT_SUCCESS;
}

// The reference's memory planner (matchUniqueImplementation.cpp:1208-1244): how many seed windows one text-side
// index block holds.  Only the order-dependent folds (scores, gaps) depend on it.  The byte counts are the reference's own,
// including their defect: AutoArray::size() returns unsigned int (AutoArray.hpp:33-36), so the size of every array is taken
// modulo 2^32 -- and the text's two arrays are even added in 32 bits (AutoTextArray.hpp:93-96) -- which on a genome makes
// the planner believe it has more room than it has.  REAL_NLIST pins the value.
namespace
{
        inline uint64_t aaSize(uint64_t n, uint64_t elt) { return (uint32_t)(sizeof(size_t) + sizeof(void *) + n * elt); }        // AutoArray<N>::size()
        inline uint64_t rankSize(uint64_t nbits)                                                                              // ERank222B::size(), ERank222B.hpp:548-555
        { return sizeof(void *) + 3 * sizeof(uint64_t) + aaSize((nbits + 65535) / 65536, 8) + aaSize((nbits + 63) / 64, 2); }
}
uint64_t planBlockWindows(RealOptions const & opts, TextFile const & T, uint64_t nreads)
{
        if ( char const * e = getenv("REAL_NLIST") ) return strtoull(e, 0, 10);
        long const pages = sysconf(_SC_PHYS_PAGES), pagesize = sysconf(_SC_PAGE_SIZE);           // = MemTotal of /proc/meminfo (getPhysicalMemory.cpp)
        uint64_t const usemem = (uint64_t)((double)((uint64_t)pages * (uint64_t)pagesize) * opts.fracmem);     // RealOptions.cpp:414-415
        uint64_t const n = T.n;
        uint64_t const textwords = ((2 * n + 63) / 64), wildwords = (n + 63) / 64;                                  // AutoTextArray.hpp:27-61
        uint64_t const textmemory = (uint64_t)(uint32_t)((uint32_t)aaSize(textwords, 8) + (uint32_t)aaSize(wildwords, 8)) + 2 * sizeof(void *) + rankSize(wildwords * 64);
        uint64_t const rvwords = (n + 1 + 63) / 64;                                                                // RangeVector.hpp:18-21,27,49
        uint64_t const rangevectormemory = aaSize(rvwords, 8) + sizeof(void *) + rankSize(rvwords * 64);
        uint64_t const lookupmemory = 2ULL * 6 * (1ULL << 22) * sizeof(size_t);                                    // 2 * getNumLists() * getHistSize() * sizeof(size_t)
        uint64_t const uniqueinfomemory = aaSize(nreads, opts.scores ? 16 : 8);                                    // AutoArray< UniqueMatchInfo<scores> >(numpat).size()
        uint64_t const nonlist = textmemory + rangevectormemory + lookupmemory + uniqueinfomemory;
        if ( nonlist > usemem )
        {
                std::cerr << "Insufficient memory." << std::endl;
                throw std::bad_alloc();
        }
        uint64_t const elementmemory = opts.seedl <= 32 ? 72 : 112;       // 3 BaseMask + (3 + getRadixSortTemp()) Mask: 3*8 + 4*12 (u32 signatures), 3*16 + 4*16 (u64)
        uint64_t const n_list_max = (usemem - nonlist) / elementmemory;
        if ( ! n_list_max )
                throw std::bad_alloc();
        uint64_t const expblocks = (T.n - opts.seedl + 1 + (n_list_max - 1)) / n_list_max;
        std::cerr << "Expected number of blocks = " << expblocks << std::endl;
        uint64_t const n_list = (T.n + (expblocks - 1)) / expblocks;
        std::cerr << "Using n_list = " << n_list << std::endl;
        return n_list;
}

int doMatchingUnique(RealOptions const & opts)
{
        PhaseTimer PT;
        GpuTeam team(opts, opts.scores || opts.gaps);     // the CUDA contexts come up while the pattern file is read
        bool const devfmt = deviceFormat();
        ReadSet reads;
        uint64_t nreads = 0;
        std::vector<std::string> filenames;
        if ( deviceReadsLoader(opts, devfmt) )
        {
                // the pattern file goes to the device as it is: parsed, packed, put into the rewritten order (-R 1) and indexed there;
                // the ids stay on the device for the formatter
                FileBytes buf;
                patternBytes(opts, buf);
                PT.lap("read patterns");
                getFileList(opts.textfilename, filenames, ".fa");
                team.wait();
                PT.lap("  contexts");
                team.shard();
                nreads = team.setReadsFasta(buf, opts.rewritepatterns);
                PT.lap("  set_reads_fasta");
                std::cerr << "Number of patterns is " << nreads << std::endl;
                team.connectFold(nreads);
        }
        else if ( opts.rewritten_input && devfmt && ! getenv("REAL_KEEP_REWRITTEN") )
        {
                // a kept rewritten pattern file: its 2 bit/base sections go to the device as they stand
                FileBytes buf;
                buf.open(opts.patternfilename);
                PackedReads packed; std::vector<uint8_t> quality; std::vector<char> idbytes; std::vector<uint64_t> idoff; bool fq = false;
                readRewrittenPacked(buf, packed, quality, idbytes, idoff, fq);
                nreads = packed.lengths.size();
                std::cerr << "Number of patterns is " << nreads << std::endl;
                PT.lap("read patterns");
                getFileList(opts.textfilename, filenames, ".fa");
                team.wait();
                team.connect(nreads, true);
                team.setReadsPacked(packed, quality);
                unsigned int const n = team.size();
                uint64_t const R = nreads;
                parallelFor(n, [&team, &idbytes, &idoff, n, R](unsigned int i) { setReadIdsBlob(team.g[i], idbytes, idoff, R * i / n, R * (i + 1) / n); });
        }
        else
        {
                loadReads(opts, reads);
                PT.lap("read patterns");
                if ( opts.rewritepatterns && ! opts.rewritten_input )
                        reorderLikeRewrite(reads);
                if ( char const * keep = getenv("REAL_KEEP_REWRITTEN") )
                {
                        // the reference writes this file for every -R 1 run and deletes it afterwards (real.cpp:238-311); kept, it can be
                        // given back as -p and spares the next run the parsing of the pattern file
                        if ( opts.rewritepatterns || opts.rewritten_input )
                        {
                                std::vector<char> bytes;
                                writeRewritten(reads, opts.fastq, bytes);
                                Output keepf(keep);
                                if ( ! bytes.empty() && fwrite(&bytes[0], 1, bytes.size(), keepf.f) != bytes.size() ) throw std::runtime_error("write failed");
                        }
                }
                PackedReads packed;
                packReads(reads, packed, hostThreads(opts));
                getFileList(opts.textfilename, filenames, ".fa");
                nreads = reads.size();
                PT.lap("  pack reads");
                team.wait();
                PT.lap("  contexts");
                team.connect(nreads, true);
                team.setReads(reads, packed);
                PT.lap("  set_reads_packed");
                if ( devfmt )
                {
                        // every handle formats the lines of the reads whose merged state it holds after the fold
                        unsigned int const n = team.size();
                        uint64_t const R = nreads;
                        parallelFor(n, [&team, &reads, n, R](unsigned int i) { setReadIds(team.g[i], reads, R * i / n, R * (i + 1) / n); });
                }
        }
        PT.lap("create + set_reads");
        Gpu & G = team.g[0];
        std::vector< std::vector< std::pair<std::string, uint64_t> > > rangeset(filenames.size());       // RangeSet
        for ( size_t fi = 0; fi < filenames.size(); ++fi )
        {
                TextFile T;
                bool const usable = setTextTeam(team, opts, (uint32_t)fi, filenames[fi], T, true);
                rangeset[fi] = T.ranges;
                if ( devfmt && fi < 64 )
                        for ( unsigned int i = 0; i < team.size(); ++i ) setRecordNames(team.g[i], (uint32_t)fi, T.ranges);
                if ( fi >= 64 || T.n >= (1ULL << 35) || T.ranges.size() > 65536 )
                {
                        std::cerr << "Skipping file " << filenames[fi] << " as it exceeds the limits of UniqueMatchInfo." << std::endl;
                        continue;
                }
                if ( ! usable )
                        continue;
                if ( opts.scores )
                        G.check(real_gpu_set_block_windows(G.h, planBlockWindows(opts, T, nreads)), "set_block_windows");
                parallelFor(team.size(), [&team](unsigned int i) { team.g[i].check(real_gpu_match_unique(team.g[i].h), "match_unique"); });
                team.foldUnique();          // several handles: every one now holds the merged state of its own share of the reads
        }
        if ( opts.gaps )
        {
                // second pass over all files for the reads still unmatched (matchUniqueImplementation.cpp:1302-1436);
                // like the reference the results are computed but not printed (the stock CLI has no output for them)
                for ( size_t fi = 0; fi < filenames.size(); ++fi )
                {
                        TextFile T;
                        if ( ! setText(G, opts, (uint32_t)fi, filenames[fi], T, true) || T.ranges.size() > 65536 )
                                continue;
                        G.check(real_gpu_match_gaps(G.h, planBlockWindows(opts, T, nreads)), "match_gaps");
                }
        }
        if ( devfmt )
        {
                PT.lap("texts + matching");
                Output out(opts.outputfilename);
                unsigned int const n = team.size();
                uint64_t const R = nreads;
                uint64_t unique = 0, written = 0;
                for ( unsigned int i = 0; i < n; ++i )
                {
                        Gpu & Gi = team.g[i];
                        uint64_t const hi = R * (i + 1) / n;
                        uint64_t at = R * i / n;
                        uint64_t const per = 1u << 18;
                        writeBatches(out, written, [&Gi, &at, &unique, hi, per](char const ** bytes, uint64_t * nbytes) -> bool
                        {
                                if ( at >= hi ) return false;
                                uint64_t const c = std::min<uint64_t>(per, hi - at);
                                uint64_t nl = 0;
                                Gi.check(real_gpu_format_unique(Gi.h, at, c, bytes, nbytes, &nl), "format_unique");
                                unique += nl; at += c;
                                return true;
                        });
                }
                PT.lap("format + write");
                std::cerr << "unique: " << unique << std::endl;
                return EXIT_SUCCESS;
        }
        std::vector<uint64_t> info(nreads + 1);
        std::vector<float> score(nreads + 1);
        {
                unsigned int const n = team.size();
                uint64_t const R = nreads;
                parallelFor(n, [&team, &info, &score, &opts, n, R](unsigned int i)
                {
                        uint64_t const lo = R * i / n, hi = R * (i + 1) / n;          // the reads whose merged state handle i holds
                        team.g[i].check(real_gpu_get_unique_range(team.g[i].h, lo, hi - lo, &info[lo], opts.scores ? &score[lo] : 0), "get_unique");
                });
        }

        PT.lap("texts + matching");
        Output out(opts.outputfilename);
        bool const scores = opts.scores;
        uint64_t const unique = formatParallel(nreads, hostThreads(opts), out,
                [&reads, &info, &score, &rangeset, scores](std::string & o, uint64_t r) -> uint64_t
        {
                uint64_t const d = info[r];
                unsigned int const state = (unsigned int)(d >> 61);
                if ( state != 1 && state != 2 )
                        return 0;
                unsigned int const file = (unsigned int)((d >> 35) & 63), frag = (unsigned int)((d >> 45) & 0xFFFF), k = (unsigned int)((d >> 41) & 15);
                uint64_t const pos = d & ((1ULL << 35) - 1);
                formatLine(o, reads, r, state == 2, scores, score[r], rangeset[file][frag].first, pos - rangeset[file][frag].second + 1, k);
                return 1;
        });
        PT.lap("format + write");
        std::cerr << "unique: " << unique << std::endl;
        return EXIT_SUCCESS;
}

int realMain(int argc, char * argv[])
{
        std::cerr << "This is REAL (B200 matching path) for the command line of REAL version 0.0.31" << std::endl;
        std::cerr << "real is distributed under version 3.0 of the GNU GENERAL PUBLIC LICENSE." << std::endl;
        try
        {
                RealOptions opts(argc, argv);
                return opts.match_unique ? doMatchingUnique(opts) : doMatchingAll(opts);
        }
        catch ( std::exception const & ex )
        {
                std::cerr << ex.what() << std::endl;
                return EXIT_FAILURE;
        }
        catch ( ... )
        {
                std::cerr << "Caught unexpected exception, terminating." << std::endl;
                return EXIT_FAILURE;
        }
}

}
