// lulesh2.0-compatible executable: see driver.cc
#include "../../../include/lulesh_host.h"
int main(int argc, char **argv) { return lulesh_host_main(argc, argv); }
