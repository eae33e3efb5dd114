// ref_harness hemi — IISPTdIntegrator::RenderView oracle (filled in with the cfg-4 work).
#include <cstdio>
int cmdHemi(int, char **) {
    fprintf(stderr, "ref_harness hemi: not built yet\n");
    return 2;
}
