/*
 * squid_b200 host twin — C entry points of the host-side stages that north_star keeps on the CPU:
 * decoding alignments into aligned blocks and loading the chimeric reads.
 *   sqh_open_case   = BuildRefName (src/ReadRec.cpp:267-283) + BuildChimericSBamRecord
 *                     (src/ReadRec.cpp:329-413) + the per-record ReadRec_t decode
 *                     (src/ReadRec.cpp:10-88) and tag / ChimName probes (src/SegmentGraph.cpp:297-302)
 *                     that the reference repeats inside each of its three BAM passes.
 * Inputs are BAM files (sqh_open_bam_case) or SQMB files (include/sqmb_format.h), the uncompressed BAM stand-in that
 * also feeds the test oracle's BamReader shim.
 */
#ifndef SQUID_B200_HOST_H
#define SQUID_B200_HOST_H
#include <stdint.h>
#include "squid_b200.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct sqh_case sqh_case;
typedef struct sqh_options {   /* src/Config.cpp:18-28 */
    int32_t phred33;           /* Phred_Type: 1 => offset 33 (Config.cpp:19) */
    int32_t max_lowphred_len;  /* -pl */
    int32_t min_phred;         /* -pm */
    int32_t min_mapq;          /* -mq; < 0 => STAR default 255 (Config.cpp:221-222) */
    int32_t concord_dist_pos, concord_dist_idx;
} sqh_options;
void sqh_default_options(sqh_options *o);
int sqh_open_case(const char *concordant_sqmb, const char *chimeric_sqmb, const sqh_options *opt, sqh_case **out, char *errbuf, int errlen);
/* The same from coordinate-sorted BAM files (BGZF-compressed or plain): replaces BamTools' BamReader::Open / GetHeader /
 * GetNextAlignment on the path (src/ReadRec.cpp:271-279, 340-343; src/SegmentGraph.cpp:293-296, 1570-1577, 3126-3129).
 * The file is inflated on all cores and decoded once for all three phases. */
int sqh_open_bam_case(const char *concordant_bam, const char *chimeric_bam, const sqh_options *opt, sqh_case **out, char *errbuf, int errlen);
void sqh_close_case(sqh_case *c);
const sqg_batch *sqh_case_batch(const sqh_case *c);
sqg_chimeric *sqh_case_chimeric(sqh_case *c);
const sqg_config *sqh_case_config(const sqh_case *c);  /* read_len filled from the chimeric reads */
int32_t sqh_case_n_ref(const sqh_case *c);
const int32_t *sqh_case_ref_len(const sqh_case *c);
/* Known-answer access to the decoder: blocks of record r of the concordant table as
 * (ref_pos, match_ref, read_pos, match_read) rows; returns the block count. */
int32_t sqh_case_blocks(const sqh_case *c, int64_t r, int32_t *out4, int32_t max_blocks, int32_t *total_len, int32_t *lowphred_run);
#ifdef __cplusplus
}
#endif
#endif
