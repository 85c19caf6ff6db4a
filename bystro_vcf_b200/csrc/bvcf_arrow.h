/* bvcf_arrow.h -- Arrow IPC file writer for the dosage matrix (bvcf_arrow.cpp), the reference's
 * bystroArrow.NewArrowIPCFileWriter + ArrowRowBuilder (main.go:320-336,517,583; arrow/arrow.go:24-137). */
#ifndef BVCF_ARROW_H
#define BVCF_ARROW_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct bvcf_arrow_writer bvcf_arrow_writer;

/* Schema: locus utf8 + one int8 column per sample, non-nullable; zstd buffer compression.  NULL on failure (err). */
bvcf_arrow_writer *bvcf_arrow_open(const char *path, const char *const *sample_names, uint32_t n_samples, char *err,
                                   size_t err_cap);
/* Append rows in order (one bvcf_dosage_batch): row-major int8 [n_rows x n_samples], loci + n_rows + 1 offsets.
 * Batches of 5,000 rows are written as they fill. */
int bvcf_arrow_write(bvcf_arrow_writer *w, uint64_t n_rows, const int8_t *dosage, const uint8_t *loci,
                     const uint64_t *loci_off);
/* Write the last (short) batch and the file footer; frees w. */
int bvcf_arrow_close(bvcf_arrow_writer *w);
const char *bvcf_arrow_error(const bvcf_arrow_writer *w);

#ifdef __cplusplus
}
#endif
#endif
