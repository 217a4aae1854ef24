// TEST INFRASTRUCTURE ONLY -- not part of the shipped product path.
//
// ref_harness <mode> <dump> [gapdump] -- <REAL command line options>
//   mode = all | unique | kat
// The options after "--" are parsed by the reference's own RealOptions
// (RealOptions.cpp:122), so flags mean exactly what they mean for src/real.
// Timing of the hot path (index build per text block, OpenMP matching region) is printed
// as one JSON line on stdout; bench.py's cpu_baseline / --impl reference legs read it.
#include "real_config.hpp"
#include "RealOptions.hpp"
#include "harness_common.hpp"

#include <iostream>
#include <stdexcept>
#include <cstring>
#include <inttypes.h>
#if defined(_OPENMP)
#include <omp.h>
#endif

int main(int argc, char * argv[])
{
        try
        {
                if ( argc < 4 )
                {
                        std::cerr << "usage: " << argv[0] << " all|unique|kat <dump> [gapdump] -- <real options>" << std::endl;
                        return 2;
                }
                std::string const mode = argv[1];
                std::string const dump = argv[2];
                std::string gapdump;
                int sep = 3;
                if ( strcmp(argv[sep],"--") != 0 )
                {
                        gapdump = argv[sep];
                        ++sep;
                }
                if ( sep >= argc || strcmp(argv[sep],"--") != 0 )
                {
                        std::cerr << "missing -- before the REAL options" << std::endl;
                        return 2;
                }

                // RealOptions skips argv[0]
                RealOptions const opts(argc - sep, argv + sep);

                HarnessTimes times;
                double const t0 = harnessNow();
                int r = 0;
                if ( mode == "all" )
                        r = harnessRunAll(opts, dump, times);
                else if ( mode == "unique" )
                        r = harnessRunUnique(opts, dump, gapdump, times);
                else if ( mode == "kat" )
                        r = harnessRunKat(opts, dump);
                else
                {
                        std::cerr << "unknown mode " << mode << std::endl;
                        return 2;
                }
                double const t1 = harnessNow();

                int threads = 1;
                #if defined(_OPENMP)
                threads = omp_get_max_threads();
                #endif
                std::cout << "{\"mode\":\"" << mode << "\",\"threads\":" << threads
                        << ",\"sort_threads\":" << opts.sort_threads
                        << ",\"reads\":" << times.reads
                        << ",\"textlen\":" << times.textlen
                        << ",\"blocks\":" << times.blocks
                        << ",\"load_s\":" << times.load_s
                        << ",\"index_s\":" << times.index_s
                        << ",\"match_s\":" << times.match_s
                        << ",\"total_s\":" << (t1-t0)
                        << "}" << std::endl;
                return r;
        }
        catch(std::exception const & ex)
        {
                std::cerr << "ref_harness: " << ex.what() << std::endl;
                return 1;
        }
}
