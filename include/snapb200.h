/*
 * snapb200.h -- C ABI of the B200-native alignment core for SNAP-RNA (andrewmagis/snap-rnaseq).
 *
 * This is the drop-in boundary.  Every entry point is `extern "C"`, takes plain
 * pointers and sizes, and replaces one reference interface, cited as file:line relative to the reference
 * tree.  The reference host code (AlignerContext, SingleAligner/PairedAligner loops, FASTQ/SAM/BAM I/O,
 * AlignmentFilter) stays as it is; an `AlignerExtension` (SNAPLib/AlignerContext.h:132-163) that owns the
 * per-thread loop drains reads into batches and calls these functions -- see INTEGRATION.md.
 *
 * Conventions
 *   - every function returns SNAPB200_OK (0) or a negative error code and never aborts the process
 *     (the reference's error model is fprintf(stderr)+soft_exit(1), SNAPLib/exit.h:26; the shim maps a
 *     non-zero return to soft_exit(1)); snapb200_last_error() gives the message;
 *   - "not aligned" is a value, not an error: status NotFound / location 0xffffffff (SNAPLib/Genome.h:29);
 *   - the caller owns all host buffers; the library owns all device memory;
 *   - there is no CPU fallback: if no CUDA device is usable every compute entry point fails with
 *     SNAPB200_ERR_CUDA.
 */
#ifndef SNAPB200_H
#define SNAPB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SNAPB200_ABI_VERSION 3

enum {
    SNAPB200_OK = 0,
    SNAPB200_ERR_ARG = -1,     /* bad argument / unsupported parameter combination           */
    SNAPB200_ERR_IO = -2,      /* index directory unreadable or malformed                     */
    SNAPB200_ERR_CUDA = -3,    /* no device, launch failure, out of device memory             */
    SNAPB200_ERR_LIMIT = -4    /* a reference pool limit was hit (the reference would soft_exit) */
};

/* AlignmentResult, SNAPLib/Read.h:41 */
enum { SNAPB200_NOT_FOUND = 0, SNAPB200_SINGLE_HIT = 1, SNAPB200_MULTIPLE_HITS = 2 };
/* Direction, SNAPLib/directions.h:26-31 */
enum { SNAPB200_FORWARD = 0, SNAPB200_RC = 1 };

#define SNAPB200_INVALID_LOCATION 0xffffffffu /* InvalidGenomeLocation, SNAPLib/Genome.h:29 */
#define SNAPB200_MAX_K 31                     /* MAX_K, SNAPLib/LandauVishkin.h:9 */
#define SNAPB200_MAX_READ_LENGTH 500          /* MAX_READ_LENGTH, SNAPLib/Read.h:45 */
#define SNAPB200_UNUSED_SCORE 0xffff          /* BaseAligner::UnusedScoreValue, SNAPLib/BaseAligner.h:261 */

typedef struct snapb200_index snapb200_index; /* opaque; a GenomeIndex + Genome resident in HBM */

/* ---- index / genome ------------------------------------------------------------------------------- */

typedef struct {
    uint32_t n_bases;             /* Genome::getCountOfBases()                                  */
    uint32_t n_pieces;            /* Genome::getNumPieces()                                     */
    uint32_t seed_len;            /* GenomeIndex::getSeedLength()                               */
    uint32_t n_hash_tables;       /* 4^(seedLen-16), SNAPLib/GenomeIndex.cpp:316                */
    uint32_t overflow_table_size; /* in 32-bit words                                            */
    uint32_t chromosome_padding;
    uint64_t hash_table_entries;  /* sum of tableSize over all tables (12 B each)               */
    uint64_t device_bytes;        /* HBM held by this handle                                    */
    int32_t device;
} snapb200_index_info;

/* Replaces GenomeIndex::loadFromDirectory (SNAPLib/GenomeIndex.cpp:844-963) + Genome::loadFromFile
 * (SNAPLib/Genome.cpp:161-261): reads `GenomeIndex`, `GenomeIndexHash`, `OverflowTable`, `Genome` from
 * `dir` and places them in the HBM of CUDA device `device`. */
int snapb200_index_open(const char *dir, int device, snapb200_index **out);

/* Same, from host memory already laid out like the files (used by tests and by hosts that keep the
 * reference's in-memory index): `tables` = concatenation of n_hash_tables tables of 12-byte entries
 * {key,value1,value2} (SNAPLib/HashTable.h:119-123), `table_sizes[i]` entries each; `overflow` =
 * OverflowTable words; `bases` = n_bases genome bytes (1 byte/base, 'n' padding included);
 * piece_offsets = Piece::beginningOffset of each piece (SNAPLib/Genome.h:155-158). */
int snapb200_index_from_memory(int device, uint32_t seed_len, uint32_t chromosome_padding,
                               uint32_t n_hash_tables, const uint64_t *table_sizes, const void *tables,
                               const uint32_t *overflow, uint32_t overflow_words,
                               const uint8_t *bases, uint32_t n_bases,
                               const uint32_t *piece_offsets, uint32_t n_pieces,
                               snapb200_index **out);

/* Device-side index construction: replaces GenomeIndex::BuildIndexToDirectory (SNAPLib/GenomeIndex.cpp:348-720)
 * with a sort-based build whose lookupSeed results are identical.  `bases` is the genome in the reference's
 * in-memory layout (1 byte/base, `chromosome_padding` 'n' before every piece and after the last one,
 * SNAPLib/FASTA.cpp:68-126); piece_names may be NULL.  slack: extra table capacity, reference default 0.3. */
int snapb200_index_build(int device, const uint8_t *bases, uint32_t n_bases, const uint32_t *piece_offsets,
                         const char *const *piece_names, uint32_t n_pieces, uint32_t seed_len,
                         uint32_t chromosome_padding, double slack, snapb200_index **out);

/* Writes the four files of the reference's index directory (formats: SNAPLib/GenomeIndex.cpp:646-710,
 * SNAPLib/HashTable.cpp:181-215, SNAPLib/Genome.cpp:126-158) so the reference can load this index. */
int snapb200_index_save(snapb200_index *idx, const char *dir);

int snapb200_index_info_get(const snapb200_index *idx, snapb200_index_info *info);
void snapb200_index_close(snapb200_index *idx);

/* ---- read batches --------------------------------------------------------------------------------- */

/* n clipped reads in ASCII (Read::getData/getQuality/getDataLength, SNAPLib/Read.h:335-339), upper case
 * (Read::init upper-cases, SNAPLib/Read.h:307-327).  Read i occupies bases[offsets[i] .. offsets[i+1]). */
typedef struct {
    uint32_t n;
    const uint32_t *offsets; /* n+1 entries, offsets[0] == 0 */
    const uint8_t *bases;
    const uint8_t *quals;
} snapb200_read_batch;

/* ---- single-end: BaseAligner ---------------------------------------------------------------------- */

/* Constructor + setter arguments of BaseAligner (SNAPLib/BaseAligner.cpp:46-61, BaseAligner.h:138-142). */
typedef struct {
    uint32_t max_hits;             /* maxHitsToConsider (-h)                                     */
    uint32_t max_k;                /* maxK (-d)                                                  */
    uint32_t max_read_size;        /* maxReadSize (MAX_READ_LENGTH in both run loops)            */
    uint32_t num_seeds;            /* maxSeedsToUseFromCommandLine (-n); 0 => use seed_coverage  */
    double seed_coverage;          /* maxSeedCoverage (-sc)                                      */
    uint32_t extra_search_depth;   /* -D                                                         */
    uint32_t explore_popular_seeds; /* -x                                                        */
    uint32_t stop_on_first_hit;    /* -f                                                         */
    uint32_t max_hits_to_get;      /* AlignRead's maxHitsToGet (0 => no multi-hit capture)       */
} snapb200_single_params;

/* Out-parameters of BaseAligner::AlignRead (SNAPLib/BaseAligner.cpp:510-524) plus the two probabilities
 * computeMAPQ consumed and the per-read counters (BaseAligner.h:167-171), for checking. */
typedef struct {
    uint32_t location;  /* *genomeLocation (0xffffffff when none)                               */
    int32_t score;      /* *finalScore (0xffff when unused)                                     */
    int32_t mapq;       /* *mapq (0 where the reference leaves it unwritten)                    */
    uint8_t status;     /* return value: NotFound/SingleHit/MultipleHits                        */
    uint8_t direction;  /* *hitDirection                                                        */
    uint16_t popular_seeds_skipped;
    uint32_t n_lookups; /* lookupSeed calls for this read (nHashTableLookups delta)             */
    uint32_t n_scored;  /* candidate locations LV-scored (nLocationsScored delta)               */
    double p_all;       /* probabilityOfAllCandidates at exit                                   */
    double p_best;      /* probabilityOfBestCandidate at exit                                   */
} snapb200_single_result;

/* Replaces BaseAligner::AlignRead, 5-argument form (SNAPLib/BaseAligner.cpp:196-200), for a batch. */
int snapb200_single_batch(snapb200_index *idx, const snapb200_single_params *params,
                          const snapb200_read_batch *reads, snapb200_single_result *results);

/* Replaces the 13-argument AlignRead with maxHitsToGet > 0 (SNAPLib/BaseAligner.cpp:510-524, 940-975;
 * caller SNAPLib/PairedAligner.cpp:584-614).  hit_* are [n][params->max_hits_to_get]; hit_counts[i] =
 * *multiHitsFound. */
int snapb200_single_multihit_batch(snapb200_index *idx, const snapb200_single_params *params,
                                   const snapb200_read_batch *reads, snapb200_single_result *results,
                                   int32_t *hit_counts, uint32_t *hit_locations, uint8_t *hit_rcs,
                                   int32_t *hit_scores);

/* Replaces BaseAligner::CharacterizeSeeds (SNAPLib/BaseAligner.cpp:206-508; callers
 * AlignmentFilter::UnalignedRead / FindPartialMatches, SNAPLib/AlignmentFilter.cpp:758, 968-971) for a batch:
 * the seed schedule and lookups of AlignRead without scoring.  The reference fills two
 * std::map<unsigned location, std::set<unsigned seedOffset>> per read (forward, RC); here segment
 * s = 2*i + direction holds that map's (location, seed offset) tuples in ascending (location, seed offset)
 * order -- the in-order traversal of the reference's containers -- at positions
 * [seg_offsets[s], seg_offsets[s+1]) of locations[] / seed_offsets[].  seg_offsets has 2*n+1 entries and is
 * always filled, so seg_offsets[2*n] is the number of tuples.  locations/seed_offsets may both be NULL
 * (count only); otherwise `capacity` is their length in tuples and SNAPB200_ERR_ARG is returned if it is too
 * small (2 * max_hits * seeds-per-read tuples per read always suffice).  Of `params`, max_hits, max_k,
 * num_seeds, seed_coverage and explore_popular_seeds are used (the partialAligner of
 * SNAPLib/PairedAligner.cpp:518-527 is max_hits 300, num_seeds 12). */
int snapb200_characterize_batch(snapb200_index *idx, const snapb200_single_params *params,
                                const snapb200_read_batch *reads, uint64_t *seg_offsets,
                                uint32_t *locations, uint16_t *seed_offsets, uint64_t capacity);

/* ---- paired-end: ChimericPairedEndAligner over IntersectingPairedEndAligner ------------------------- */

/* Constructor arguments of IntersectingPairedEndAligner (SNAPLib/IntersectingPairedEndAligner.cpp:34-49)
 * and ChimericPairedEndAligner (SNAPLib/ChimericPairedEndAligner.cpp:41-61). */
typedef struct {
    uint32_t max_hits;            /* maxHits of the single-end fallback BaseAligner (-h)         */
    uint32_t max_k;               /* -d                                                          */
    uint32_t max_read_size;
    uint32_t num_seeds;           /* -n (0 => seed_coverage)                                     */
    double seed_coverage;
    uint32_t min_spacing;         /* -s min                                                      */
    uint32_t max_spacing;         /* -s max                                                      */
    uint32_t force_spacing;       /* -fs                                                         */
    uint32_t max_big_hits;        /* intersectingAlignerMaxHits (-H)                             */
    uint32_t extra_search_depth;  /* -D                                                          */
    uint32_t max_candidate_pool_size; /* -mcp                                                    */
} snapb200_paired_params;

/* PairedAlignmentResult (SNAPLib/PairedEndAligner.h:31-56) fields the aligners fill, plus the pair
 * probabilities fed to computeMAPQ (IntersectingPairedEndAligner.cpp:741). */
typedef struct {
    uint32_t location[2];
    int32_t score[2];
    int32_t mapq[2];
    uint8_t status[2];
    uint8_t direction[2];
    uint8_t from_align_together;
    uint8_t aligned_as_pair;
    uint16_t pad;
    uint32_t n_lv_calls; /* scoreLocation calls in the intersecting aligner for this pair          */
    uint32_t n_lookups;  /* lookupSeed calls in the intersecting aligner for this pair             */
    double p_all;        /* probabilityOfAllPairs  (0 when the intersecting aligner returned early) */
    double p_best;       /* probabilityOfBestPair                                                  */
} snapb200_paired_result;

/* Replaces ChimericPairedEndAligner::align (SNAPLib/ChimericPairedEndAligner.cpp:74-128), i.e.
 * IntersectingPairedEndAligner::align (SNAPLib/IntersectingPairedEndAligner.cpp:141-753) plus the
 * per-end BaseAligner fallback with mapq/4.  reads0->n must equal reads1->n. */
int snapb200_paired_batch(snapb200_index *idx, const snapb200_paired_params *params,
                          const snapb200_read_batch *reads0, const snapb200_read_batch *reads1,
                          snapb200_paired_result *results);

/* ---- CIGAR: LandauVishkinWithCigar at SAM-write time ------------------------------------------------ */

/* Replaces SAMFormat::computeCigarString's aligner call (SNAPLib/SAM.cpp:1159-1189) for a batch: read i
 * (clipped, as given to the aligner; reverse-complemented on the device when directions[i]==RC, as
 * SAM.cpp:870-880 does) is re-aligned against genome[locations[i], +len) with k = MAX_K-1 and the
 * COMPACT_CIGAR_STRING written NUL-terminated to cigars + i*cigar_stride.  edit_distance[i] is the return
 * value of LandauVishkinWithCigar::computeEditDistance (-1: > k, -2: buffer too small, -3: location invalid
 * or off the end of the genome, where the reference prints "*").  Soft-clip decoration stays on the host. */
int snapb200_cigar_batch(snapb200_index *idx, const snapb200_read_batch *reads, const uint32_t *locations,
                         const uint8_t *directions, int use_m, char *cigars, uint32_t cigar_stride,
                         int32_t *edit_distance);

/* ---- the I/O edges of the path as batch kernels ---------------------------- */

/* n reads as the reference's readers hand them to the run loops: Read::getId/getIdLength, the UNCLIPPED bases and
 * qualities (Read::getUnclippedData/getUnclippedQuality/getUnclippedLength, upper-cased as Read::init does,
 * SNAPLib/Read.h:295-327) and the clipping Read::clip chose (getFrontClippedLength, getDataLength;
 * SNAPLib/Read.h:357-404).  Read i: bases/quals[offsets[i] .. offsets[i+1]), ids[id_offsets[i] .. id_offsets[i+1]). */
typedef struct {
    uint32_t n;
    const uint32_t *offsets;     /* n+1 */
    const uint8_t *bases;
    const uint8_t *quals;
    const uint16_t *front_clip;  /* n */
    const uint16_t *clipped_len; /* n */
    const uint32_t *id_offsets;  /* n+1 */
    const uint8_t *ids;
} snapb200_sam_reads;

/* Replaces FASTQReader::getNextRead (SNAPLib/FASTQ.cpp:188-246: four lines per record, CR LF tolerated, the
 * starting-character table of :253-297) + Read::init (upper-casing) + Read::clip(clipping) for every complete record
 * of `text` (which must begin at a record start, as after FASTQReader::reinit).  clipping: ReadClippingType
 * (SNAPLib/Read.h:85: 0 none, 1 front, 2 back, 3 both).  Outputs are the arrays of a snapb200_sam_reads (each of
 * offsets/id_offsets has max_reads+1 entries; bases, quals and ids must hold n_bytes bytes, which always suffices);
 * *bytes_consumed = offset of the first byte after the last complete record (what data->advance() has consumed).
 * The clipped read the aligners take is bases + offsets[i] + front_clip[i], clipped_len[i] bytes.
 * A blank line or an invalid starting character -- where the reference prints and soft_exit(1)s -- returns
 * SNAPB200_ERR_ARG with the reference's message in snapb200_last_error(); more than max_reads records or a read
 * longer than 65535 bases: SNAPB200_ERR_LIMIT. */
int snapb200_fastq_parse(int device, const uint8_t *text, uint64_t n_bytes, int clipping, uint32_t max_reads,
                         uint32_t *n_reads, uint64_t *bytes_consumed, uint32_t *offsets, uint8_t *bases, uint8_t *quals,
                         uint16_t *front_clip, uint16_t *clipped_len, uint32_t *id_offsets, uint8_t *ids);

/* Replaces FASTQReader::skipPartialRecord (SNAPLib/FASTQ.cpp:113-184), which FASTQReader::reinit runs when a reader is given a
 * byte range that does not start at offset 0 (RangeSplitter hands every worker thread such ranges, SNAPLib/RangeSplitter.h:37-55):
 * *offset = where the first record begins in `text` (n_bytes if there is none), i.e. where the text given to
 * snapb200_fastq_parse must start.  Host logic (it looks at a few lines; no device is involved), which is how one file is
 * sharded over host threads and GPUs. */
int snapb200_fastq_record_start(const uint8_t *text, uint64_t n_bytes, uint64_t *offset);

/* The arguments of ReadWriter::writeRead / the per-end fields of PairedAlignmentResult that SAM output uses
 * (SNAPLib/Read.h:171, SNAPLib/PairedEndAligner.h:31-56).  skip != 0: emit nothing for this read.  is_transcriptome != 0: the
 * alignment was found in the transcriptome at tlocation (location is its genome position, as AlignmentFilter reports it); only
 * snapb200_sam_batch_rna takes such alignments. */
typedef struct {
    uint32_t location;
    int32_t mapq;
    uint8_t status;    /* AlignmentResult */
    uint8_t direction; /* Direction */
    uint8_t skip;
    uint8_t is_transcriptome; /* PairedAlignmentResult::isTranscriptome / writeRead's isTranscriptome */
    uint32_t tlocation;       /* PairedAlignmentResult::tlocation */
} snapb200_sam_alignment;

/* use_m of the entry points below: 0, or SNAPB200_SAM_USE_M for AlignerOptions::useM (-M: M instead of = and X); with
 * SNAPB200_SAM_BAM_RECORDS or-ed in, every "line" is the BAM record BAMFormat::writeRead puts into the writer's buffer instead
 * (SNAPLib/Bam.cpp:596-790: the BAMAlignment head with bin / n_cigar_op / l_seq, NUL-terminated name -- not cut at a space --, binary
 * CIGAR operations, 4-bit bases, qualities minus '!', RG:Z / PG:Z:SNAP / NM:i), uncompressed: BGZF stays the writer's filter.  NM of
 * a read without a location is an uninitialised variable in the reference (Bam.cpp:644); -1 is written.  A name of more than 254
 * bytes (the reference exits, Bam.cpp:723): SNAPB200_ERR_LIMIT. */
#define SNAPB200_SAM_USE_M 1
#define SNAPB200_SAM_BAM_RECORDS 2

/* Replaces SimpleReadWriter::writeRead / writePair (SNAPLib/ReadWriter.cpp:90-217) over SAMFormat::writeRead
 * (SNAPLib/SAM.cpp:803-1153: getSAMData, computeCigarString with LandauVishkinWithCigar at k = MAX_K-1, soft clips,
 * FLAG/RNEXT/PNEXT/TLEN, the /1 /2 QNAME trimming, "\tPG:Z:SNAP\tNM:i:%d") for genome alignments of a batch.
 * reads1/aln1 == NULL: single-end, line i = read i.  Otherwise pair p yields lines 2p and 2p+1 in the order
 * writePair writes them (the end with the lower location first; it also gets SAM_FIRST_SEGMENT, as in the
 * reference).  Lines are concatenated into out (no header); line_offsets has n_lines+1 entries and is always
 * filled.  out == NULL: only measure.  out_capacity too small: SNAPB200_ERR_ARG (line_offsets is valid, so the
 * caller knows the size).  read_group: ReaderContext::defaultReadGroup or NULL.  Bytes are identical to the
 * reference's for reads without auxiliary data (FASTQ input). */
int snapb200_sam_batch(snapb200_index *idx, const snapb200_sam_reads *reads0, const snapb200_sam_reads *reads1,
                       const snapb200_sam_alignment *aln0, const snapb200_sam_alignment *aln1, int use_m,
                       const char *read_group, char *out, uint64_t out_capacity, uint64_t *line_offsets);

/* (opened by snapb200_annotation_open, below) */
typedef struct snapb200_annotation snapb200_annotation; /* exon / gene tables of a GTF in HBM, for a genome + transcriptome index pair */

/* The same for RNA mode: alignments with is_transcriptome set get their CIGAR from the transcriptome text at tlocation with the
 * splice junctions of their transcript inserted (SAMFormat::writeRead's transcriptome branch, SNAPLib/SAM.cpp:1046-1061, over
 * LandauVishkinWithCigar::insertSpliceJunctions, SNAPLib/LandauVishkin.cpp:119-250, and GTFTranscript::Junctions,
 * SNAPLib/GTFReader.cpp:1109-1139).  genome and transcriptome must be the pair the annotation was opened with and live on its
 * device.  A spliced CIGAR longer than 511 characters: SNAPB200_ERR_LIMIT. */
int snapb200_sam_batch_rna(snapb200_annotation *annotation, snapb200_index *genome, snapb200_index *transcriptome,
                           const snapb200_sam_reads *reads0, const snapb200_sam_reads *reads1, const snapb200_sam_alignment *aln0,
                           const snapb200_sam_alignment *aln1, int use_m, const char *read_group, char *out, uint64_t out_capacity,
                           uint64_t *line_offsets);

/* BGZF (SAM/BAM specification section 4.1): the container GzipWriterFilter::compressChunk builds for BAM output
 * (SNAPLib/GzipDataWriter.cpp:281-340: one gzip member per chunk with the "BC" extra field holding the member's size, then CRC-32 and
 * ISIZE).  Replaces it for the CONTENT, not for zlib's bytes: every `chunk` input bytes (0: 65024, the maximum) become one member whose
 * deflate stream is this library's own -- a dynamic-Huffman block over the chunk's byte histogram, no string matching, or a stored
 * block when that is not smaller -- so any BGZF reader inflates exactly `data`, but the file is neither zlib's bytes nor as small as
 * zlib's.  n_bytes == 0 gives one empty member.  out_capacity must be at least n_bytes + 31 * ceil(n_bytes / chunk) (the stored worst
 * case); block_offsets (optional) receives n_blocks + 1 offsets into out. */
int snapb200_bgzf_compress(int device, const uint8_t *data, uint64_t n_bytes, uint32_t chunk, uint8_t *out, uint64_t out_capacity,
                           uint64_t *out_bytes, uint64_t *block_offsets);
int snapb200_bgzf_last_kernel_ms(float *ms);

/* CUDA-event times (ms) of the kernels of the last snapb200_fastq_parse / snapb200_sam_batch call made by this thread
 * (no copies), for the streaming roofline of these two stages. */
int snapb200_io_last_kernel_ms(float *fastq_ms, float *sam_ms);

/* ---- RNA mode: AlignmentFilter on the device ------------------------------
 * One warp per pair (snap_rnaseq_b200/csrc/filter_warp.cuh) over the per-element rules of csrc/filterfmt.h, which are verified on
 * the host against the reference's AlignmentFilter.  The annotation is loaded with the reference's GTFReader semantics
 * (csrc/gtf_tables.h). */

/* Replaces GTFReader::Load (SNAPLib/GTFReader.cpp:1245-1361) for what the filter reads.  Every transcriptome piece must be a
 * transcript of the annotation and every transcript's chromosome a piece of the genome (the reference exits otherwise).  The handle
 * keeps its own copies of what it needs from the two indices (piece tables, names), so it may outlive them; it lives on the genome
 * index's device. */
int snapb200_annotation_open(snapb200_index *genome, snapb200_index *transcriptome, const char *gtf_path, snapb200_annotation **out);
void snapb200_annotation_close(snapb200_annotation *a);
/* Names behind the indices of snapb200_filter_event: transcript ids in the order of the reference's transcript map (std::map by id),
 * chromosome names = genome piece names.  NULL when out of range.  Valid until snapb200_annotation_close. */
uint32_t snapb200_annotation_transcript_count(const snapb200_annotation *a);
const char *snapb200_annotation_transcript_id(const snapb200_annotation *a, uint32_t index);
const char *snapb200_annotation_chromosome(const snapb200_annotation *a, uint32_t index);

typedef struct {
    uint32_t max_spacing;      /* -s max                                         */
    uint32_t force_spacing;    /* -fs                                            */
    uint32_t conf_diff;        /* AlignerOptions::confDiff (-c)                   */
    uint32_t max_dist;         /* options->maxDist.start (-d)                     */
    uint32_t max_hits_to_get;  /* row stride of the multi-hit arrays (1000, SNAPLib/PairedAligner.cpp:584) */
} snapb200_filter_params;

/* The fields of PairedAlignmentResult that leave the loop after AlignmentFilter::Filter, forceSpacing and the MAPQ halving
 * (SNAPLib/PairedAligner.cpp:620-663).  fromAlignTogether is always false after Filter. */
typedef struct {
    uint32_t location[2], tlocation[2];
    int32_t score[2], mapq[2];
    uint8_t status[2], direction[2], is_transcriptome[2];
    uint8_t aligned_as_pair; /* result->alignedAsPair: set only when a same-gene pair decided (AlignmentFilter.cpp:543-548) */
    uint8_t pad;
} snapb200_filter_result;

/* What the host still has to count for the pair, through the reference's own public GTFReader methods:
 * kind 1 IncrementReadCount, 2 IntrachromosomalPair, 3 InterchromosomalPair with the two chosen alignments (transcript = index in
 * transcript-id order or -1, chr = genome piece index); unaligned 1 / 2: AlignmentFilter::UnalignedRead(read 0 / read 1). */
typedef struct {
    int32_t kind, unaligned, transcript[2], chr[2];
    uint32_t pos_original[2], pos[2], pos_end[2];
} snapb200_filter_event;

/* One novel-splice candidate of AlignmentFilter::UnalignedRead (SNAPLib/AlignmentFilter.cpp:742-933): the two partial alignments of a
 * read without any alignment that the reference hands to GTFReader::IntrachromosomalSplice (kind 2) or InterchromosomalSplice (kind 3)
 * together with the read's id (read 0's or read 1's of pair `pair`, as snapb200_filter_event::unaligned says).  chr = genome piece. */
typedef struct {
    uint32_t pair;
    int32_t kind;
    int32_t chr[2];
    uint32_t pos[2], pos_end[2];
} snapb200_splice;

/* Replaces the AlignmentFilter section of PairedAlignerContext::runIterationThread (SNAPLib/PairedAligner.cpp:582-663;
 * AlignmentFilter::AddAlignment / Filter, SNAPLib/AlignmentFilter.cpp:140-740) for a batch of n pairs.  Inputs are what the other
 * entry points return: the transcriptome multi-hits of both reads (snapb200_single_multihit_batch: counts[n] and rows of
 * max_hits_to_get), the genome pairs (snapb200_paired_batch) and the CharacterizeSeeds tuples of both reads
 * (snapb200_characterize_batch with the partialAligner's parameters: seg_offsets[2n+1], locations, seed_offsets).  len0/len1: data
 * lengths of the reads.  needs_host[i] != 0: the pair has more alignments or combinations than the device scratch holds; its
 * result and event are not valid and the caller runs the reference's AlignmentFilter for it. */
int snapb200_filter_paired_batch(snapb200_annotation *a, const snapb200_filter_params *params, uint32_t n, const uint32_t *len0,
                                 const uint32_t *len1, const int32_t *n0, const uint32_t *loc0, const uint8_t *rc0, const int32_t *score0,
                                 const int32_t *n1, const uint32_t *loc1, const uint8_t *rc1, const int32_t *score1,
                                 const snapb200_paired_result *genome_pairs, const uint64_t *seg0, const uint32_t *ch_loc0,
                                 const uint16_t *ch_off0, const uint64_t *seg1, const uint32_t *ch_loc1, const uint16_t *ch_off1,
                                 snapb200_filter_result *results, snapb200_filter_event *events, uint8_t *needs_host);

/* ---- RNA mode, whole pair loop: what one iteration of PairedAlignerContext::runIterationThread computes for a batch ------------
 * (SNAPLib/PairedAligner.cpp:582-663): transcriptomeAligner->AlignRead x2 with maxHitsToGet, g_aligner->align, the partialAligner's
 * CharacterizeSeeds for both reads, AlignmentFilter::Filter, forceSpacing and the MAPQ halving -- with every intermediate (multi-hit
 * lists, genome pairs, seed tuples) staying in HBM.  A batch object owns pinned host staging for its inputs and outputs and one
 * worker thread: submit() copies the reads into pinned memory and returns; wait() blocks until the device work and the downloads are
 * done.  Two objects per host thread give the double buffering north_star asks for: the host replays batch k (GTF counters, SAM
 * output) while batch k+1 is on the device.  The three handles must live on the same device. */
typedef struct snapb200_rna_batch snapb200_rna_batch;

typedef struct {
    snapb200_paired_params paired;         /* g_aligner (ChimericPairedEndAligner), PairedAligner.cpp:470-481               */
    snapb200_single_params transcriptome;  /* transcriptomeAligner incl. max_hits_to_get (1000), PairedAligner.cpp:512, 584 */
    snapb200_single_params partial;        /* partialAligner: max_hits 300, num_seeds 12, PairedAligner.cpp:518-527        */
    snapb200_filter_params filter;
} snapb200_rna_params;

/* Pointers into the batch object's pinned host memory; valid until the next submit on the same object.
 * hits: CSR over reads -- read i of mate e has hit_offsets[e][i+1] - hit_offsets[e][i] transcriptome multi-hits (AlignRead's
 * multiHitLocations / RCs / Scores in the reference's order).  seg_offsets / ch_*: snapb200_characterize_batch layout. */
typedef struct {
    uint32_t n;
    const snapb200_filter_result *results;
    const snapb200_filter_event *events;
    const uint8_t *needs_host;
    const snapb200_paired_result *genome_pairs;
    const uint32_t *hit_offsets[2];
    const uint32_t *hit_locations[2];
    const uint8_t *hit_rcs[2];
    const int32_t *hit_scores[2];
    const uint64_t *seg_offsets[2];
    const uint32_t *ch_locations[2];
    const uint16_t *ch_seed_offsets[2];
    /* AlignmentFilter::UnalignedRead of every read events[i].unaligned flags, as records in the reference's loop order: pair i owns
     * splices[splice_offsets[i] .. splice_offsets[i+1]).  splice_overflow[i] != 0: the read has more partial alignments than the
     * device scratch holds; no records, the caller runs UnalignedRead itself (from seg_offsets / ch_*). */
    const uint64_t *splice_offsets;
    const snapb200_splice *splices;
    const uint8_t *splice_overflow;
    float device_ms;   /* wall time of the device work of this batch (uploads, kernels, downloads), for the shim's timing report */
    /* snapb200_rna_batch_submit_sam only (NULL otherwise): what SimpleReadWriter::writePair writes for pair i is
     * sam_text[sam_line_offsets[2i] .. sam_line_offsets[2i+2]) -- two lines, the end with the lower location first -- with the
     * filter's result (after the forceSpacing rule of PairedAligner.cpp:648-651, applied when params.paired.force_spacing or
     * params.filter.force_spacing is set) as the alignment.  A pair whose range is empty
     * was not formatted (needs_host[i], or a spliced CIGAR too long for its slot): the caller writes it with the reference's writer. */
    const char *sam_text;
    const uint64_t *sam_line_offsets;
} snapb200_rna_view;

int snapb200_rna_batch_create(snapb200_annotation *a, snapb200_index *genome, snapb200_index *transcriptome, snapb200_rna_batch **out);
void snapb200_rna_batch_destroy(snapb200_rna_batch *b);
int snapb200_rna_batch_submit(snapb200_rna_batch *b, const snapb200_rna_params *params, const snapb200_read_batch *reads0,
                              const snapb200_read_batch *reads1);
/* The same, and the SAM lines of every pair as the last stage of the submission (snapb200_sam_batch_rna on the resident reads and
 * results): sam0 / sam1 hold the unclipped reads, their ids and the clipping of the same n pairs (reads0 / reads1 are their clipped
 * parts, as the aligners see them). */
int snapb200_rna_batch_submit_sam(snapb200_rna_batch *b, const snapb200_rna_params *params, const snapb200_read_batch *reads0,
                                  const snapb200_read_batch *reads1, const snapb200_sam_reads *sam0, const snapb200_sam_reads *sam1,
                                  int use_m, const char *read_group);
int snapb200_rna_batch_wait(snapb200_rna_batch *b, snapb200_rna_view *view);

/* ---- building blocks exposed for known-answer tests -------------------------------------------------- */

/* LandauVishkin<+1/-1>::computeEditDistance (SNAPLib/LandauVishkin.h:211-455) on explicit strings.
 * Item i: text = texts[text_offsets[i] .. text_offsets[i+1]), pattern/quality likewise.  For
 * text_direction == -1 the text is given in memory order and is walked backwards from its END (the
 * reference is handed a pointer one past the last byte).  Bytes outside either string never match.
 * quals may be NULL (then match_probability is not produced).  Outputs: score (-1 if > k),
 * match_probability, net_indel. */
int snapb200_lv_batch(int device, int text_direction, uint32_t n, const uint32_t *text_offsets,
                      const uint8_t *texts, const uint32_t *pattern_offsets, const uint8_t *patterns,
                      const uint8_t *quals, const int32_t *k, int32_t *score, double *match_probability,
                      int32_t *net_indel);

/* LandauVishkinWithCigar::computeEditDistance (SNAPLib/LandauVishkin.cpp:252-535) on explicit strings,
 * COMPACT_CIGAR_STRING format. */
int snapb200_lv_cigar_batch(int device, uint32_t n, const uint32_t *text_offsets, const uint8_t *texts,
                            const uint32_t *pattern_offsets, const uint8_t *patterns, const int32_t *k,
                            int use_m, char *cigars, uint32_t cigar_stride, int32_t *edit_distance);

/* GenomeIndex::lookupSeed (SNAPLib/GenomeIndex.cpp:971-1011) for n seeds given as ASCII (seed_len bytes
 * each, contiguous).  n_hits is [n][2] (forward, reverse complement); up to max_out hits per direction are
 * copied to hits[(i*2+dir)*max_out ..].  A seed containing a non-ACGT byte yields 0/0. */
int snapb200_lookup_seed_batch(snapb200_index *idx, uint32_t n, const uint8_t *seeds, uint32_t max_out,
                               uint32_t *n_hits, uint32_t *hits);

/* Diagnostics for the roofline of stage 2 (index probes) in isolation.
 * snapb200_probe_bench: n seeds taken from the resident genome at `positions[i]` (seed_len bases each; positions with a
 *   non-ACGT base count as a lookup of nothing) are packed and looked up exactly as the aligners do -- one lane per seed,
 *   both directions resolved, overflow count words read -- `iters` times; returns the CUDA-event time of one pass and the
 *   totals of one pass: table slots examined (12 B each), overflow count words read (4 B each), hits reported.
 * snapb200_gather_bench: the ceiling to compare against: `n` independent uniformly random 32-byte sector reads over a
 *   `bytes`-sized buffer, one per thread, `iters` passes; returns the time of one pass. */
int snapb200_probe_bench(snapb200_index *idx, uint32_t n, const uint32_t *positions, uint32_t iters, float *ms_per_pass,
                         uint64_t *slots_examined, uint64_t *count_words, uint64_t *hits_reported);
int snapb200_gather_bench(int device, uint64_t bytes, uint32_t n, uint32_t iters, float *ms_per_pass);

/* computeMAPQ (SNAPLib/mapq.h:32-65), evaluated on the device the way the aligner kernels do. */
int snapb200_mapq_batch(int device, uint32_t n, const double *p_all, const double *p_best,
                        const int32_t *score, const int32_t *popular_seeds_skipped, int32_t *mapq);
/* Same; host_reevaluated[i] != 0 where the device's -10*log10(1 - pBest/pAll) landed within 1e-9 of an integer and the library
 * re-evaluated computeMAPQ with the host's libm, as the aligner entry points do (the value the reference truncates, mapq.h:51, is
 * last-ulp sensitive exactly there).  For tests that force this path. */
int snapb200_mapq_batch_ex(int device, uint32_t n, const double *p_all, const double *p_best, const int32_t *score,
                           const int32_t *popular_seeds_skipped, int32_t *mapq, uint8_t *host_reevaluated);

/* ---- device-resident sessions (what bench.py times; also what a pipelined host uses) ---------------- */

typedef struct snapb200_session snapb200_session; /* device buffers + stream for batches of <= max_reads */

int snapb200_session_create(snapb200_index *idx, uint32_t max_pairs_or_reads, uint32_t max_read_len,
                            snapb200_session **out);
void snapb200_session_destroy(snapb200_session *s);
/* async H2D of a batch into mate slot 0/1 (slot 0 for single-end) from (preferably pinned) host memory */
int snapb200_session_upload(snapb200_session *s, int slot, const snapb200_read_batch *reads);
/* run the kernels on the resident batch; no host<->device copies */
int snapb200_session_run_single(snapb200_session *s, const snapb200_single_params *p);
int snapb200_session_run_paired(snapb200_session *s, const snapb200_paired_params *p);
/* async D2H of the results of the last run, then stream synchronize */
int snapb200_session_download_single(snapb200_session *s, snapb200_single_result *results);
int snapb200_session_download_paired(snapb200_session *s, snapb200_paired_result *results);
int snapb200_session_sync(snapb200_session *s);
/* CUDA-event time of the last run's kernels (ms), kernel launches issued by the last run, and the number
 * of launches since the session was created. */
int snapb200_session_last_run(const snapb200_session *s, float *kernel_ms, uint32_t *launches,
                              uint64_t *total_launches);
/* CUDA-event time (ms) of the dominant kernel of the last run alone: the first single_kernel / paired_kernel
 * launch (small scratch tier, all items).  Used for the roofline figure. */
int snapb200_session_main_kernel_ms(const snapb200_session *s, float *main_ms);

/* ---- statistics (AlignerStats, SNAPLib/AlignerStats.h:45-60) ------------------------------------------ */

typedef struct {
    int64_t total_reads, useful_reads, single_hits, multi_hits, not_found, errors, aligned_as_pairs, lv_calls;
    int64_t n_hash_table_lookups, n_locations_scored, n_hits_ignored_popularity, n_reads_ignored_ns;
    int64_t n_table_probes;   /* 12-byte hash-table slots examined by lookupSeed (>= 1 per lookup)          */
    int64_t n_hit_words_read; /* 4-byte hit-list words the aligners consumed (votes, binary-search probes) */
    int64_t mapq_histogram[71];
} snapb200_stats;
#define SNAPB200_STATS_WORDS (14 + 71)

/* Accumulated over every batch run on this index handle since the last reset; a flat int64 vector so that
 * the multi-GPU host can sum it with one all-reduce (NCCL), as AlignerStats::add does per thread
 * (SNAPLib/AlignerStats.cpp:75-102). */
int snapb200_stats_get(snapb200_index *idx, snapb200_stats *out);
/* The sum over the per-GPU copies of one index (one host process drives all GPUs of a box): what
 * AlignerStats::add does over the reference's per-thread objects. */
int snapb200_stats_sum(snapb200_index *const *indices, uint32_t n, snapb200_stats *out);
int snapb200_stats_reset(snapb200_index *idx);

const char *snapb200_last_error(void);
int snapb200_abi_version(void);
int snapb200_device_count(void);

#ifdef __cplusplus
}
#endif
#endif /* SNAPB200_H */
