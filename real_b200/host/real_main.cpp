// `real` command line on the GPU matching path (real.cpp:357-375).  Like the reference's main() the
// process exit code is 0 also after a reported error; set REAL_STRICT_EXIT=1 to get the failure code.
#include "real_host.hpp"
#include <cstdlib>

int main(int argc, char * argv[])
{
        int const rc = realhost::realMain(argc, argv);
        return getenv("REAL_STRICT_EXIT") ? rc : 0;
}
