/*
 * fastq_b200.h -- host feed of libbbduk_b200.so: FASTQ text <-> the batch layout of bbduk_b200_process
 * (SURVEY.md 8f row 1). Plain C, host only (no GPU work): once the k-mer block runs at G reads/s the wall clock
 * is the parse/format of the text, so both are native, multithreaded and copy each byte once.
 *
 * Replaces, for uncompressed 4-line FASTQ: the record splitting of stream/FastqStreamer.java +
 * stream/FASTQ.java (quad parsing: header line '@', bases, '+' line, qualities; "\r\n" tolerated) and the record
 * formatting of stream/ReadStreamByteWriter.java for the reads the k-mer block keeps
 * (routing as jgi/BBDuk.java:3190-3254: pairs that are not removed go to out, trimmed to [lo,hi); removed pairs go
 * to outm). kmask / ksplit rewrite bases and are formatted by the caller.
 *
 * rec[] holds 4 int64 per read: {header_start, bases_start, bases_len, quals_start}, positions inside the text
 * buffer the read came from (header_start points at '@'). Returns 0 on success, non-zero on a malformed record.
 */
#ifndef FASTQ_B200_H
#define FASTQ_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define FASTQ_API __attribute__((visibility("default")))
#else
#define FASTQ_API
#endif

/* Index the complete records of text[0..n_bytes). Record r is written to rec[4*(first + r*stride)] (stride 2 interleaves
 * the mates of two files: file 1 with first=0, file 2 with first=1). final!=0: the buffer ends the file, so a last line
 * without '\n' is complete. *consumed = bytes covered by the indexed records (the caller keeps the rest for the next
 * block). At most max_records are indexed. rec == NULL: only count (*n_records = complete 4-line records). */
FASTQ_API int fastq_b200_index(const uint8_t *text, int64_t n_bytes, int32_t final, int64_t max_records, int64_t stride,
                               int64_t first, int64_t *rec, int64_t *n_records, int64_t *consumed, int32_t threads);

/* offsets[0..n_reads] = running sum of bases_len; bases[] = the reads' bases, concatenated (the layout of
 * bbduk_b200_process). Read i comes from text2 if text2!=NULL and i is odd, else from text1. */
FASTQ_API int fastq_b200_gather(const uint8_t *text1, const uint8_t *text2, const int64_t *rec, int64_t n_reads,
                                uint8_t *bases, int64_t *offsets, int32_t threads);

/* Format the kept (want_removed=0) or removed (want_removed!=0) units as FASTQ text. per = 1 (single reads) or 2
 * (pairs, mates adjacent); mate_sel: 0 = every mate (interleaved output), 1 = first mates only, 2 = second mates only.
 * Kept reads are cut to [lo,hi); removed reads are written untrimmed unless trim_removed!=0 (ottm).
 * out may be NULL to query the size; *out_len = bytes needed / written. */
FASTQ_API int fastq_b200_format(const uint8_t *text1, const uint8_t *text2, const int64_t *rec, int64_t n_reads, int32_t per,
                                const int32_t *lo, const int32_t *hi, const uint8_t *flags, int32_t want_removed,
                                int32_t mate_sel, int32_t trim_removed, uint8_t *out, int64_t out_cap, int64_t *out_len,
                                int32_t threads);

#ifdef __cplusplus
}
#endif
#endif /* FASTQ_B200_H */
