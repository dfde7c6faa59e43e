// snap-rna-b200: the reference's command line with the alignment core on the GPU.
//
// Same sub-commands and options as apps/snap/Main.cpp:54-84 of the reference; the only difference is that the
// single/paired contexts are constructed with a GpuAlignerExtension (the two-line change INTEGRATION.md shows).
// `index` and `transcriptome` are the reference's own host-side builders.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "stdafx.h"
#include "GenomeIndex.h"
#include "PairedAligner.h"
#include "SingleAligner.h"
#include "exit.h"

#include "GpuAlignerExtension.h"

static void reportExit()
{
    if (getenv("SNAPB200_SHIM_TIMING") != NULL)
        fprintf(stderr, "[snapb200 shim] process exits %.2f s after its start (the handlers registered later, the CUDA runtime's among them, have run)\n",
                GpuAlignerExtension::sinceProcessStart());
}

int main(int argc, const char **argv)
{
    GpuAlignerExtension::sinceProcessStart();
    atexit(reportExit);
    const char *version = "0.1alpha-b200";
    printf("Welcome to SNAP-RNA version %s.\n\n", version);
    if (argc < 2) {
        fprintf(stderr, "Usage: snap-rna-b200 <index|transcriptome|single|paired> [<options>]\n");
        soft_exit(1);
    } else if (strcmp(argv[1], "index") == 0) {
        GenomeIndex::runIndexer(argc - 2, argv + 2);
    } else if (strcmp(argv[1], "transcriptome") == 0) {
        GenomeIndex::runTranscriptomeIndexer(argc - 2, argv + 2);
    } else {
        for (int i = 1; i < argc;) {
            unsigned nArgsConsumed = 0;
            if (strcmp(argv[i], "single") == 0 || strcmp(argv[i], "paired") == 0) {
                // the hash tables live in HBM only: the host keeps the genome text (SAM output) and the seed length
                setenv("SNAPB200_GENOME_ONLY", "1", 1);
                // SAM lines that come back formatted from the device are placed by the reference's writer as they are (RNA pair loop)
                PreformattedSAMFormat::install();
                // optional: HBM loading overlaps the host-side loading
                GpuAlignerExtension::prefetch(argc - (i + 1), argv + i + 1, strcmp(argv[i], "paired") == 0);
            }
            if (strcmp(argv[i], "single") == 0) {
                SingleAlignerContext single(new GpuAlignerExtension());
                single.runAlignment(argc - (i + 1), argv + i + 1, version, &nArgsConsumed);
            } else if (strcmp(argv[i], "paired") == 0) {
                PairedAlignerContext paired(new GpuAlignerExtension());
                paired.runAlignment(argc - (i + 1), argv + i + 1, version, &nArgsConsumed);
                if (getenv("SNAPB200_SHIM_TIMING") != NULL)
                    fprintf(stderr, "[snapb200 shim] runAlignment returned %.2f s after process start (the reference's end-of-run GTF analysis included)\n",
                            GpuAlignerExtension::sinceProcessStart());
            } else {
                fprintf(stderr, "Invalid command: %s\n", argv[i]);
                soft_exit(1);
            }
            i += nArgsConsumed + 1;
        }
    }
    return 0;
}
