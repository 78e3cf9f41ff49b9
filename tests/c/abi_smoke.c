/* Plain-C client of include/peppan_b200.h: proves the boundary is usable without C++, Python or torch.  Built and run by
 * tests/test_abi.py.  Only entry points that need no GPU do work here; pb_init must fail loudly on a box without a device. */
#include <stdio.h>
#include <string.h>
#include "peppan_b200.h"

int main(void)
{
    /* one hit: query ACGTACGTAC against the same bases at 3..12 of the contig, CIGAR 10M */
    const char* q = "ACGTACGTAC"; const char* t = "TTACGTACGTACTT";
    int64_t qoff[2] = {0, 10}, toff[2] = {0, 14};
    pb_seqset qs = {(const uint8_t*)q, qoff, 1}, ts = {(const uint8_t*)t, toff, 1};
    int32_t qid = 0, sid = 0, qstart = 1, qend = 10, sstart = 3, send = 12;
    int64_t coff[2] = {0, 1};
    uint32_t cig[1] = {(10u << 2) | 0u};
    double iden = 0, score = 0;
    int rc = pb_rescore_m1(&qs, &ts, 1, &qid, &sid, &qstart, &qend, &sstart, &send, coff, cig, &iden, &score);
    if (rc != PB_OK || iden != 1.0 || score != 30.0) { printf("rescore failed rc=%d iden=%g score=%g\n", rc, iden, score); return 1; }

    pb_post_params prm; memset(&prm, 0, sizeof(prm));
    prm.fix_start = 3; prm.fix_end = 3; prm.do_overlap = 1; prm.ovl_len = 300; prm.ovl_prop = 0.6;
    int32_t qrank = 0, srank = 0, qlen = 10, slen = 14, hid = 7;
    pb_post_result res;
    rc = pb_post_chain(1, &qrank, &srank, &iden, &score, &qstart, &qend, &sstart, &send, &qlen, &slen, &hid, coff, cig, &prm, &res);
    if (rc != PB_OK || res.n_rows != 1 || res.row[0] != 0 || res.n_overlaps != 0) { printf("post chain failed rc=%d\n", rc); return 2; }
    pb_free_post(&res);

    pb_ctx* ctx = NULL;
    rc = pb_init(0, 0, 1, NULL, &ctx);
    if (rc == PB_OK) { printf("gpu present: sm count query\n"); int32_t sm = 0, khz = 0; int64_t mem = 0; pb_device_info(ctx, &sm, &khz, &mem); pb_destroy(ctx); printf("C ABI OK (gpu, %d SMs)\n", sm); return 0; }
    if (rc != PB_ERR_NODEVICE || strstr(pb_last_error(NULL), "no CPU fallback") == NULL) { printf("unexpected pb_init result %d: %s\n", rc, pb_last_error(NULL)); return 3; }
    printf("C ABI OK (no device: %s)\n", pb_last_error(NULL));
    return 0;
}
