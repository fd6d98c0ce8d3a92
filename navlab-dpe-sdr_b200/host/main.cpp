// main.cpp -- dpe_console: the reference's `cudarecv` entry point for this path
// (cudarecv/cudarecv/src/main.cu:37-84): an interactive / scripted shell over FlowMgr.
//   dpe_console                 commands from stdin
//   dpe_console -f script.do    run a dofile, then exit
#include <csignal>
#include <cstring>
#include <fstream>
#include <iostream>
#include "console.h"

static dsp::FlowMgr* g_mgr = nullptr;
static void onSigint(int) { if (g_mgr) g_mgr->EmergencyStop(); std::_Exit(130); }

int main(int argc, char** argv) {
    dsp::FlowMgr mgr;
    g_mgr = &mgr;
    std::signal(SIGINT, onSigint);
    console::Shell shell(&mgr);
    if (argc >= 3 && std::strcmp(argv[1], "-f") == 0) {
        std::ifstream f(argv[2]);
        if (!f) { std::cerr << "cannot open " << argv[2] << std::endl; return 2; }
        const int rc = shell.run(f, false);
        mgr.EmergencyStop();
        return rc < 0 ? 1 : 0;
    }
    const int rc = shell.run(std::cin, true);
    mgr.EmergencyStop();
    return rc < 0 ? 1 : 0;
}
