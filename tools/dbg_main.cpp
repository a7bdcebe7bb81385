#include <cstdio>
#include <cuda_runtime.h>
#include "../include/circulantpc.h"
int main() {
    fprintf(stderr, "devices %d\n", cpc_device_count());
    cpc_plan_desc d = {4, 1, 1, 1, 0, 1, 0, nullptr, nullptr, -1};
    cpc_plan p = nullptr;
    int rc = cpc_plan_create(&p, &d);
    fprintf(stderr, "create rc=%d %s\n", rc, cpc_last_error());
    if (rc) return 1;
    rc = cpc_set_symbol_transport(p, 1.0, 0, 0);
    fprintf(stderr, "symbol rc=%d %s\n", rc, cpc_last_error());
    double h[8] = {0, 0, 1, 0, 8, 0, 27, 0}, o[8];
    rc = cpc_apply(p, h, o, CPC_MEM_HOST);
    fprintf(stderr, "apply rc=%d %s\n", rc, cpc_last_error());
    for (int i = 0; i < 4; ++i) fprintf(stderr, "%g %g\n", o[2 * i], o[2 * i + 1]);
    cpc_destroy(p);
    return 0;
}
