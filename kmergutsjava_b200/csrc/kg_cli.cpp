// kg_cli.cpp -- kmer_guts_b200: the command line of KmerGutsJava.main (KGJ:560-654) over libkmerguts_b200.so.
#include "../../include/kmerguts_host.h"

int main(int argc, char** argv) { return kg_main(argc, argv); }
