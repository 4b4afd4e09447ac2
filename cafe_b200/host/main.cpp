// main.cpp — `cafe_gpu_shell script`: runs a CAFE command script (the likelihood-path subset of the
// reference's shell, main.cpp:24-67) on the GPU path.  Used for end-to-end lambda-hat parity runs.
#include <cstdio>
#include <cstdlib>
#include <ctime>

#include "cafe_commands.h"

int main(int argc, char* argv[]) {
    Globals globals;
    std::srand((unsigned)std::time(nullptr));
    if (argc != 2) {
        std::fprintf(stderr, "usage: %s <script>\n", argv[0]);
        return 2;
    }
    std::vector<std::string> tokens{"source", argv[1]};
    int rc = 0;
    try {
        rc = cafe_cmd_source(globals, tokens);
    } catch (std::exception& e) {
        std::fprintf(stderr, "%s\n", e.what());
        rc = -1;
    }
    cafe_gpu_engine_release();
    return rc ? 1 : 0;
}
