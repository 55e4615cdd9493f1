// wepp_main.cpp — the executable Snakemake calls (`build/wepp`, workflow/rules/filter.smk:21-36,
// workflow/rules/sam2pb.smk:28-40): argv goes straight to the library (include/wepp_b200.h).
#include "../../include/wepp_b200.h"

int main(int argc, char** argv) { return wepp_cli_main(argc, argv); }
