// ref_driver.cpp -- C API over the UNMODIFIED reference classes (TEST INFRASTRUCTURE ONLY).
//
// Linked by oracle/build_ref.py against the reference's own objects into oracle/_ref/libsnapref.so.
// It contains no alignment logic of its own: every function constructs the reference object the run
// loops construct (SNAPLib/SingleAligner.cpp:167-181, SNAPLib/PairedAligner.cpp:459-481) and calls the
// reference method, copying its out-parameters into the structs of include/snapb200.h so that the
// oracle port, the reference and the CUDA library can be compared field by field.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs load this.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <time.h>
#include <pthread.h>
#include <map>
#include <set>
#include <vector>
#include <string>
#include <algorithm>

// The reference keeps the probabilities computeMAPQ consumes in private members; the driver reads them.
#define private public
#define protected public
#include "stdafx.h"
#include "Compat.h"
#include "Genome.h"
#include "GenomeIndex.h"
#include "Seed.h"
#include "Read.h"
#include "LandauVishkin.h"
#include "BaseAligner.h"
#include "IntersectingPairedEndAligner.h"
#include "ChimericPairedEndAligner.h"
#include "mapq.h"
#include "Tables.h"
#include "BigAlloc.h"
#include "FASTQ.h"
#include "SAM.h"
#include "FileFormat.h"
#include "DataReader.h"
#include "DataWriter.h"
#include "AlignmentFilter.h"
#include "GTFReader.h"
#include "ProbabilityDistance.h"
#undef private
#undef protected

#include "../include/snapb200.h"

extern "C" {

int ref_init(void)
{
    initializeLVProbabilitiesToPhredPlus33(); // SNAPLib/AlignerOptions.cpp:84 does this for every run
    return 0;
}

void *ref_index_load(const char *dir)
{
    ref_init();
    char *d = strdup(dir);
    GenomeIndex *idx = GenomeIndex::loadFromDirectory(d);
    free(d);
    return idx;
}

int ref_index_info(void *h, snapb200_index_info *info)
{
    GenomeIndex *idx = (GenomeIndex *)h;
    memset(info, 0, sizeof(*info));
    info->n_bases = idx->getGenome()->getCountOfBases();
    info->n_pieces = idx->getGenome()->getNumPieces();
    info->seed_len = idx->getSeedLength();
    info->n_hash_tables = idx->nHashTables;
    info->overflow_table_size = idx->overflowTableSize;
    info->chromosome_padding = idx->getGenome()->chromosomePadding;
    for (unsigned i = 0; i < idx->nHashTables; i++) info->hash_table_entries += idx->hashTables[i]->GetTableSize();
    info->device = -1;
    return 0;
}

// copies genome bytes [from, from+len) (no bounds games: caller stays inside [0,nBases))
int ref_genome_bytes(void *h, unsigned from, unsigned len, unsigned char *out)
{
    GenomeIndex *idx = (GenomeIndex *)h;
    const Genome *g = idx->getGenome();
    if ((size_t)from + len > g->getCountOfBases()) return -1;
    memcpy(out, g->bases + from, len);
    return 0;
}

int ref_piece_offsets(void *h, unsigned *out)
{
    const Genome *g = ((GenomeIndex *)h)->getGenome();
    for (int i = 0; i < g->getNumPieces(); i++) out[i] = g->getPieces()[i].beginningOffset;
    return g->getNumPieces();
}

int ref_lookup_seed_batch(void *h, unsigned n, const unsigned char *seeds, unsigned max_out, unsigned *n_hits,
                          unsigned *hits)
{
    GenomeIndex *idx = (GenomeIndex *)h;
    unsigned L = idx->getSeedLength();
    for (unsigned i = 0; i < n; i++) {
        const char *s = (const char *)seeds + (size_t)i * L;
        unsigned nh[2] = {0, 0};
        const unsigned *hp[2] = {NULL, NULL};
        if (Seed::DoesTextRepresentASeed(s, L)) {
            Seed seed(s, L);
            idx->lookupSeed(seed, &nh[0], &hp[0], &nh[1], &hp[1]);
        }
        for (int d = 0; d < 2; d++) {
            n_hits[i * 2 + d] = nh[d];
            for (unsigned j = 0; j < nh[d] && j < max_out; j++) hits[((size_t)i * 2 + d) * max_out + j] = hp[d][j];
        }
    }
    return 0;
}

// Explicit-string LV.  The reference compares 8 bytes at a time and may peek past either string, so each
// string is copied into a padded buffer whose padding can never match the other string's padding.
static const int PAD = 64;

int ref_lv_batch(int text_direction, unsigned n, const unsigned *text_offsets, const unsigned char *texts,
                 const unsigned *pattern_offsets, const unsigned char *patterns, const unsigned char *quals,
                 const int *k, int *score, double *match_probability, int *net_indel)
{
    ref_init();
    LandauVishkin<1> *fwd = new LandauVishkin<1>();
    LandauVishkin<-1> *rev = new LandauVishkin<-1>();
    std::vector<char> tbuf, pbuf, qbuf;
    for (unsigned i = 0; i < n; i++) {
        int tl = text_offsets[i + 1] - text_offsets[i];
        int pl = pattern_offsets[i + 1] - pattern_offsets[i];
        tbuf.assign(tl + 2 * PAD, 1);
        pbuf.assign(pl + 2 * PAD, 0);
        qbuf.assign(pl + 2 * PAD, '!');
        memcpy(&tbuf[PAD], texts + text_offsets[i], tl);
        memcpy(&pbuf[PAD], patterns + pattern_offsets[i], pl);
        if (quals) memcpy(&qbuf[PAD], quals + pattern_offsets[i], pl);
        double prob = 0;
        int indel = 0;
        int s;
        if (text_direction == 1) {
            s = fwd->computeEditDistance(&tbuf[PAD], tl, &pbuf[PAD], quals ? &qbuf[PAD] : NULL, pl, k[i],
                                         quals ? &prob : NULL, 0, &indel);
        } else {
            s = rev->computeEditDistance(&tbuf[PAD] + tl, tl, &pbuf[PAD], quals ? &qbuf[PAD] : NULL, pl, k[i],
                                         quals ? &prob : NULL, 0, &indel);
        }
        score[i] = s;
        if (match_probability) match_probability[i] = prob;
        if (net_indel) net_indel[i] = indel;
    }
    delete fwd;
    delete rev;
    return 0;
}

int ref_lv_cigar_batch(unsigned n, const unsigned *text_offsets, const unsigned char *texts,
                       const unsigned *pattern_offsets, const unsigned char *patterns, const int *k, int use_m,
                       char *cigars, unsigned cigar_stride, int *edit_distance)
{
    LandauVishkinWithCigar *lv = new LandauVishkinWithCigar();
    std::vector<char> tbuf, pbuf;
    for (unsigned i = 0; i < n; i++) {
        int tl = text_offsets[i + 1] - text_offsets[i];
        int pl = pattern_offsets[i + 1] - pattern_offsets[i];
        tbuf.assign(tl + 2 * PAD, 1);
        pbuf.assign(pl + 2 * PAD, 0);
        memcpy(&tbuf[PAD], texts + text_offsets[i], tl);
        memcpy(&pbuf[PAD], patterns + pattern_offsets[i], pl);
        std::vector<unsigned> tokens;
        char *out = cigars + (size_t)i * cigar_stride;
        memset(out, 0, cigar_stride);
        edit_distance[i] = lv->computeEditDistance(&tbuf[PAD], tl, &pbuf[PAD], pl, k[i], out, cigar_stride,
                                                   use_m != 0, tokens);
    }
    delete lv;
    return 0;
}

int ref_mapq_batch(unsigned n, const double *p_all, const double *p_best, const int *score, const int *popular,
                   int *mapq)
{
    for (unsigned i = 0; i < n; i++) mapq[i] = computeMAPQ(p_all[i], p_best[i], score[i], popular[i]);
    return 0;
}

// ProbabilityDistance::compute (SNAPLib/ProbabilityDistance.cpp:53-135); the object is 1.2 MB, hence the heap
int ref_probability_distance(double snp_prob, double gap_open_prob, double gap_extension_prob, const char *reference, const char *read,
                             const char *quality, int read_len, int max_start_shift, int max_shift, double *match_probability)
{
    ProbabilityDistance *pd = new ProbabilityDistance(snp_prob, gap_open_prob, gap_extension_prob);
    int rc = pd->compute(reference, read, quality, read_len, max_start_shift, max_shift, match_probability);
    delete pd;
    return rc;
}

// ------------------------------------------------------------------------------------------------
// Threaded batch drivers.  One aligner object per thread, exactly like the reference's run loops.
// ------------------------------------------------------------------------------------------------
struct SingleJob {
    GenomeIndex *idx;
    const snapb200_single_params *p;
    const snapb200_read_batch *reads;
    snapb200_single_result *res;
    int *hit_counts;
    unsigned *hit_locations;
    unsigned char *hit_rcs;
    int *hit_scores;
    unsigned begin, end;
};

static void *single_worker(void *arg)
{
    SingleJob *j = (SingleJob *)arg;
    const snapb200_single_params *p = j->p;
    BaseAligner *a = new BaseAligner(j->idx, p->max_hits, p->max_k, p->max_read_size, p->num_seeds, p->seed_coverage,
                                     p->extra_search_depth, NULL, NULL, NULL, NULL);
    a->setExplorePopularSeeds(p->explore_popular_seeds != 0);
    a->setStopOnFirstHit(p->stop_on_first_hit != 0);
    unsigned mh = p->max_hits_to_get;
    std::vector<char> bases, quals; // padded copies: the reference reads 8 bytes at a time past the read
    bool *rcs = mh ? new bool[mh] : NULL;
    for (unsigned i = j->begin; i < j->end; i++) {
        unsigned off = j->reads->offsets[i], len = j->reads->offsets[i + 1] - off;
        bases.assign(len + PAD, '\n');
        quals.assign(len + PAD, '\n');
        memcpy(&bases[0], j->reads->bases + off, len);
        memcpy(&quals[0], j->reads->quals + off, len);
        Read read;
        read.init(NULL, 0, &bases[0], &quals[0], len);
        snapb200_single_result *r = &j->res[i];
        memset(r, 0, sizeof(*r));
        unsigned loc = InvalidGenomeLocation;
        Direction dir = FORWARD;
        int score = 0, mapq = 0;
        _int64 l0 = a->getNHashTableLookups(), s0 = a->getLocationsScored();
        a->probabilityOfAllCandidates = 0;
        a->probabilityOfBestCandidate = 0;
        a->popularSeedsSkipped = 0;
        AlignmentResult st;
        if (mh) {
            int found = 0;
            st = a->AlignRead(&read, &loc, &dir, &score, &mapq, 0, 0, FORWARD, (int)mh, &found,
                              j->hit_locations + (size_t)i * mh, rcs, j->hit_scores + (size_t)i * mh);
            j->hit_counts[i] = found;
            for (int q = 0; q < found; q++) j->hit_rcs[(size_t)i * mh + q] = rcs[q] ? 1 : 0;
        } else {
            st = a->AlignRead(&read, &loc, &dir, &score, &mapq);
        }
        r->status = (uint8_t)st;
        r->location = loc;
        r->direction = (uint8_t)dir;
        r->score = score;
        r->mapq = mapq;
        r->popular_seeds_skipped = (uint16_t)a->popularSeedsSkipped;
        r->n_lookups = (uint32_t)(a->getNHashTableLookups() - l0);
        r->n_scored = (uint32_t)(a->getLocationsScored() - s0);
        r->p_all = a->probabilityOfAllCandidates;
        r->p_best = a->probabilityOfBestCandidate;
    }
    delete[] rcs;
    delete a;
    return NULL;
}

static int run_single(void *h, const snapb200_single_params *p, const snapb200_read_batch *reads,
                      snapb200_single_result *res, int *hc, unsigned *hl, unsigned char *hr, int *hs, int nthreads)
{
    ref_init();
    if (nthreads < 1) nthreads = 1;
    std::vector<SingleJob> jobs(nthreads);
    std::vector<pthread_t> th(nthreads);
    for (int t = 0; t < nthreads; t++) {
        SingleJob &j = jobs[t];
        j.idx = (GenomeIndex *)h; j.p = p; j.reads = reads; j.res = res;
        j.hit_counts = hc; j.hit_locations = hl; j.hit_rcs = hr; j.hit_scores = hs;
        j.begin = (unsigned)((unsigned long long)reads->n * t / nthreads);
        j.end = (unsigned)((unsigned long long)reads->n * (t + 1) / nthreads);
    }
    if (nthreads == 1) { single_worker(&jobs[0]); return 0; }
    for (int t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, single_worker, &jobs[t]);
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    return 0;
}

int ref_single_batch(void *h, const snapb200_single_params *p, const snapb200_read_batch *reads,
                     snapb200_single_result *res, int nthreads)
{
    snapb200_single_params q = *p;
    q.max_hits_to_get = 0;
    return run_single(h, &q, reads, res, NULL, NULL, NULL, NULL, nthreads);
}

int ref_single_multihit_batch(void *h, const snapb200_single_params *p, const snapb200_read_batch *reads,
                              snapb200_single_result *res, int *hit_counts, unsigned *hit_locations,
                              unsigned char *hit_rcs, int *hit_scores, int nthreads)
{
    return run_single(h, p, reads, res, hit_counts, hit_locations, hit_rcs, hit_scores, nthreads);
}

// probabilityOfAllPairs / probabilityOfBestPair of the last IntersectingPairedEndAligner::align on this thread: locals of align()
// in the reference, copied out by the two inserted lines of oracle/build_ref.py (INSERT_PATCHES)
extern __thread double snapref_pair_p[2];

// The aligner objects one worker thread of PairedAlignerContext::runIterationThread owns (SNAPLib/PairedAligner.cpp:459-481)
struct PairedAligners {
    BigAllocator *alloc;
    IntersectingPairedEndAligner *inter;
    ChimericPairedEndAligner *chim;
};

static PairedAligners *paired_aligners_create(GenomeIndex *index, const snapb200_paired_params *p)
{
    PairedAligners *a = new PairedAligners();
    size_t pool = IntersectingPairedEndAligner::getBigAllocatorReservation(
        index, p->max_big_hits, p->max_read_size, index->getSeedLength(), p->num_seeds, p->seed_coverage, p->max_k,
        p->extra_search_depth, p->max_candidate_pool_size);
    a->alloc = new BigAllocator(pool);
    a->inter = new IntersectingPairedEndAligner(
        index, p->max_read_size, p->max_hits, p->max_k, p->num_seeds, p->seed_coverage, p->min_spacing, p->max_spacing,
        p->max_big_hits, p->extra_search_depth, p->max_candidate_pool_size, a->alloc);
    a->chim = new ChimericPairedEndAligner(
        index, p->max_read_size, p->max_hits, p->max_k, p->num_seeds, p->seed_coverage, p->min_spacing, p->max_spacing,
        p->force_spacing != 0, p->extra_search_depth, a->inter);
    return a;
}

static void paired_aligners_destroy(PairedAligners *a)
{
    delete a->chim;
    a->inter->~IntersectingPairedEndAligner();
    delete a->alloc;
    delete a;
}

struct PairedJob {
    GenomeIndex *idx;
    const snapb200_paired_params *p;
    const snapb200_read_batch *r0, *r1;
    snapb200_paired_result *res;
    unsigned begin, end;
    PairedAligners *aligners;  // NULL: construct for this call and destroy afterwards
};

static void *paired_worker(void *arg)
{
    PairedJob *j = (PairedJob *)arg;
    PairedAligners *own = j->aligners ? NULL : paired_aligners_create(j->idx, j->p);
    PairedAligners *al = j->aligners ? j->aligners : own;
    IntersectingPairedEndAligner *inter = al->inter;
    ChimericPairedEndAligner *chim = al->chim;
    std::vector<char> b[2], q[2];
    for (unsigned i = j->begin; i < j->end; i++) {
        Read reads[2];
        const snapb200_read_batch *rb[2] = {j->r0, j->r1};
        for (int e = 0; e < 2; e++) {
            unsigned off = rb[e]->offsets[i], len = rb[e]->offsets[i + 1] - off;
            b[e].assign(len + PAD, '\n');
            q[e].assign(len + PAD, '\n');
            memcpy(&b[e][0], rb[e]->bases + off, len);
            memcpy(&q[e][0], rb[e]->quals + off, len);
            reads[e].init(NULL, 0, &b[e][0], &q[e][0], len);
        }
        PairedAlignmentResult pr;
        memset(&pr, 0, sizeof(pr)); // the reference leaves this uninitialised (PairedAligner.cpp:577)
        pr.location[0] = pr.location[1] = InvalidGenomeLocation;
        _int64 s0 = inter->getLocationsScored();
        inter->countOfHashTableLookups[0] = inter->countOfHashTableLookups[1] = 0;
        snapref_pair_p[0] = snapref_pair_p[1] = 0;
        chim->align(&reads[0], &reads[1], &pr);
        snapb200_paired_result *r = &j->res[i];
        memset(r, 0, sizeof(*r));
        for (int e = 0; e < 2; e++) {
            r->location[e] = pr.location[e];
            r->score[e] = pr.score[e];
            r->mapq[e] = pr.mapq[e];
            r->status[e] = (uint8_t)pr.status[e];
            r->direction[e] = (uint8_t)pr.direction[e];
        }
        r->from_align_together = pr.fromAlignTogether;
        r->aligned_as_pair = pr.alignedAsPair;
        r->n_lv_calls = (uint32_t)(inter->getLocationsScored() - s0);
        r->n_lookups = inter->countOfHashTableLookups[0] + inter->countOfHashTableLookups[1];
        r->p_all = snapref_pair_p[0];  // 0 when IntersectingPairedEndAligner::align returned before phase 3
        r->p_best = snapref_pair_p[1];
    }
    if (own) paired_aligners_destroy(own);
    return NULL;
}

int ref_paired_batch(void *h, const snapb200_paired_params *p, const snapb200_read_batch *r0,
                     const snapb200_read_batch *r1, snapb200_paired_result *res, int nthreads)
{
    ref_init();
    if (nthreads < 1) nthreads = 1;
    std::vector<PairedJob> jobs(nthreads);
    std::vector<pthread_t> th(nthreads);
    for (int t = 0; t < nthreads; t++) {
        PairedJob &j = jobs[t];
        j.idx = (GenomeIndex *)h; j.p = p; j.r0 = r0; j.r1 = r1; j.res = res; j.aligners = NULL;
        j.begin = (unsigned)((unsigned long long)r0->n * t / nthreads);
        j.end = (unsigned)((unsigned long long)r0->n * (t + 1) / nthreads);
    }
    if (nthreads == 1) { paired_worker(&jobs[0]); return 0; }
    for (int t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, paired_worker, &jobs[t]);
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    return 0;
}

// The same with the aligner objects of every thread kept alive between calls, as a worker thread of the reference keeps them for a
// whole run (SNAPLib/PairedAligner.cpp:459-527 constructs once per thread per run): what bench.py --impl reference times, so that
// constructing a BigAllocator + IntersectingPairedEndAligner + ChimericPairedEndAligner is not inside every timed step.
struct PairedPool {
    GenomeIndex *idx;
    snapb200_paired_params p;
    std::vector<PairedAligners *> aligners;
};

void *ref_paired_pool_create(void *h, const snapb200_paired_params *p, int nthreads)
{
    ref_init();
    PairedPool *pool = new PairedPool();
    pool->idx = (GenomeIndex *)h;
    pool->p = *p;
    for (int t = 0; t < (nthreads < 1 ? 1 : nthreads); t++) pool->aligners.push_back(paired_aligners_create(pool->idx, p));
    return pool;
}

int ref_paired_pool_run(void *vp, const snapb200_read_batch *r0, const snapb200_read_batch *r1, snapb200_paired_result *res)
{
    PairedPool *pool = (PairedPool *)vp;
    const int nthreads = (int)pool->aligners.size();
    std::vector<PairedJob> jobs(nthreads);
    std::vector<pthread_t> th(nthreads);
    for (int t = 0; t < nthreads; t++) {
        PairedJob &j = jobs[t];
        j.idx = pool->idx; j.p = &pool->p; j.r0 = r0; j.r1 = r1; j.res = res; j.aligners = pool->aligners[t];
        j.begin = (unsigned)((unsigned long long)r0->n * t / nthreads);
        j.end = (unsigned)((unsigned long long)r0->n * (t + 1) / nthreads);
    }
    for (int t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, paired_worker, &jobs[t]);
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    return 0;
}

void ref_paired_pool_destroy(void *vp)
{
    PairedPool *pool = (PairedPool *)vp;
    for (size_t t = 0; t < pool->aligners.size(); t++) paired_aligners_destroy(pool->aligners[t]);
    delete pool;
}

// SAMFormat::computeCigarString's aligner call (SNAPLib/SAM.cpp:1159-1189): text = genome at location,
// textLen = patternLen = read length, k = MAX_K-1.  RC reads are complemented first (SAM.cpp getSAMData).
int ref_cigar_batch(void *h, const snapb200_read_batch *reads, const unsigned *locations,
                    const unsigned char *directions, int use_m, char *cigars, unsigned cigar_stride,
                    int *edit_distance)
{
    GenomeIndex *idx = (GenomeIndex *)h;
    const Genome *genome = idx->getGenome();
    LandauVishkinWithCigar *lv = new LandauVishkinWithCigar();
    std::vector<char> pbuf;
    for (unsigned i = 0; i < reads->n; i++) {
        unsigned off = reads->offsets[i], len = reads->offsets[i + 1] - off;
        char *out = cigars + (size_t)i * cigar_stride;
        memset(out, 0, cigar_stride);
        if (locations[i] == InvalidGenomeLocation) { edit_distance[i] = -3; continue; }
        const char *ref = genome->getSubstring(locations[i], len);
        if (ref == NULL) { edit_distance[i] = -3; continue; }
        pbuf.assign(len + PAD, 0);
        if (directions[i] == RC) {
            for (unsigned q = 0; q < len; q++) pbuf[q] = COMPLEMENT[reads->bases[off + len - 1 - q]];
        } else {
            memcpy(&pbuf[0], reads->bases + off, len);
        }
        std::vector<unsigned> tokens;
        edit_distance[i] = lv->computeEditDistance(ref, len, &pbuf[0], len, MAX_K - 1, out, cigar_stride, use_m != 0,
                                                   tokens);
    }
    delete lv;
    return 0;
}

// BaseAligner::CharacterizeSeeds (SNAPLib/BaseAligner.cpp:206-508) with the partialAligner construction of
// SNAPLib/PairedAligner.cpp:518-527; the two seed_maps are flattened in iteration order.
int ref_characterize_batch(void *h, const snapb200_single_params *p, const snapb200_read_batch *reads,
                           unsigned long long *seg_offsets, unsigned *locations, unsigned short *seed_offsets,
                           unsigned long long capacity, int nthreads)
{
    (void)nthreads;
    ref_init();
    GenomeIndex *idx = (GenomeIndex *)h;
    BaseAligner *a = new BaseAligner(idx, p->max_hits, p->max_k, p->max_read_size, p->num_seeds, p->seed_coverage,
                                     p->extra_search_depth, NULL, NULL, NULL, NULL);
    a->setExplorePopularSeeds(p->explore_popular_seeds != 0);
    a->setStopOnFirstHit(p->stop_on_first_hit != 0);
    std::vector<char> bases, quals;
    unsigned long long pos = 0;
    int rc = 0;
    seg_offsets[0] = 0;
    for (unsigned i = 0; i < reads->n && !rc; i++) {
        unsigned off = reads->offsets[i], len = reads->offsets[i + 1] - off;
        bases.assign(len + PAD, '\n');
        quals.assign(len + PAD, '\n');
        memcpy(&bases[0], reads->bases + off, len);
        memcpy(&quals[0], reads->quals + off, len);
        Read read;
        read.init(NULL, 0, &bases[0], &quals[0], len);
        seed_map maps[2];
        unsigned loc = InvalidGenomeLocation;
        Direction dir = FORWARD;
        int score = 0, mapq = 0;
        a->CharacterizeSeeds(&read, &loc, &dir, &score, &mapq, 0, 0, FORWARD, maps[0], maps[1]);
        for (int d = 0; d < 2 && !rc; d++) {
            for (seed_map::iterator it = maps[d].begin(); it != maps[d].end() && !rc; ++it) {
                for (std::set<unsigned>::iterator js = it->second.begin(); js != it->second.end(); ++js) {
                    if (locations) {
                        if (pos >= capacity) { rc = SNAPB200_ERR_ARG; break; }
                        locations[pos] = it->first;
                        seed_offsets[pos] = (unsigned short)*js;
                    }
                    pos++;
                }
            }
            seg_offsets[2 * (size_t)i + d + 1] = pos;
        }
    }
    delete a;
    return rc;
}

// ---- row f2: the reference's FASTQ reader and SAM writer, driven as the run loops drive them ----------------------------

// FASTQReader::create + getNextRead until the file ends (SNAPLib/FASTQ.cpp:55-69, 188-246) with ReaderContext::clipping as
// given; every Read is copied out in the layout of snapb200_sam_reads.
static double g_last_seconds = 0;  // time spent inside the reference's own calls by the last ref_fastq_parse / ref_sam_batch
static double now_s() { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }
double ref_last_seconds(void) { return g_last_seconds; }

int ref_fastq_parse(const char *path, int clipping, unsigned max_reads, unsigned *n_reads, unsigned *offsets, unsigned char *bases,
                    unsigned char *quals, unsigned short *front_clip, unsigned short *clipped_len, unsigned *id_offsets, unsigned char *ids)
{
    ReaderContext ctx;
    memset(&ctx, 0, sizeof(ctx));
    ctx.clipping = (ReadClippingType)clipping;
    FASTQReader *fq = FASTQReader::create(DataSupplier::Default[false], path, 0, 0, ctx);
    unsigned n = 0;
    offsets[0] = 0;
    id_offsets[0] = 0;
    Read read;
    g_last_seconds = 0;
    for (;;) {
        double t0 = now_s();
        bool more = n < max_reads && fq->getNextRead(&read);
        g_last_seconds += now_s() - t0;
        if (!more) break;
        unsigned o = offsets[n], io = id_offsets[n], len = read.getUnclippedLength();
        memcpy(bases + o, read.getUnclippedData(), len);
        memcpy(quals + o, read.getUnclippedQuality(), len);
        memcpy(ids + io, read.getId(), read.getIdLength());
        front_clip[n] = (unsigned short)read.getFrontClippedLength();
        clipped_len[n] = (unsigned short)read.getDataLength();
        offsets[n + 1] = o + len;
        id_offsets[n + 1] = io + read.getIdLength();
        n++;
    }
    *n_reads = n;
    delete fq;
    return 0;
}

// FASTQReader::create with a starting offset runs reinit -> skipPartialRecord (SNAPLib/FASTQ.cpp:88-112); *file_offset = where
// the reader stands afterwards (the first record it will return), or -1 when it found none.
int ref_fastq_record_start(const char *path, long long start, long long *file_offset)
{
    ReaderContext ctx;
    memset(&ctx, 0, sizeof(ctx));
    FASTQReader *fq = FASTQReader::create(DataSupplier::Default[false], path, start, 0, ctx);
    char *buffer;
    _int64 valid;
    if (!fq->data->getData(&buffer, &valid) || valid <= 0 || buffer[0] != '@') *file_offset = -1;
    else *file_offset = fq->data->getFileOffset();
    delete fq;
    return 0;
}

static void make_read(Read *r, const snapb200_sam_reads *b, unsigned i, const char *read_group)
{
    unsigned o = b->offsets[i], len = b->offsets[i + 1] - o;
    r->init((const char *)b->ids + b->id_offsets[i], b->id_offsets[i + 1] - b->id_offsets[i], (const char *)b->bases + o,
            (const char *)b->quals + o, len);
    // the state Read::clip leaves (Read.h:357-404)
    r->data += b->front_clip[i];
    r->quality += b->front_clip[i];
    r->frontClippedLength = b->front_clip[i];
    r->dataLength = b->clipped_len[i];
    r->setReadGroup(read_group);
}

// SimpleReadWriter::writeRead / writePair (SNAPLib/ReadWriter.cpp:90-217) through the reference's own ReadWriterSupplier and
// file DataWriter into `path` (no header).  Reads with skip set are not written.
int ref_sam_batch_rna(void *h, void *h_transcriptome, void *gtf, const snapb200_sam_reads *r0, const snapb200_sam_reads *r1,
                      const snapb200_sam_alignment *a0, const snapb200_sam_alignment *a1, int use_m, const char *read_group, const char *path)
{
    GenomeIndex *idx = (GenomeIndex *)h;
    const Genome *genome = idx->getGenome();
    const Genome *transcriptome = h_transcriptome ? ((GenomeIndex *)h_transcriptome)->getGenome() : NULL;
    DataWriterSupplier *dws = DataWriterSupplier::create(path);
    // use_m: bit 0 = AlignerOptions::useM, bit 1 = BAM records (BAMFormat::writeRead) instead of SAM lines -- straight into the file,
    // without the gzip filter BAMFormat::getWriterSupplier would add (Bam.cpp:515-536)
    const FileFormat *format = (use_m & 2) ? FileFormat::BAM[use_m & 1] : FileFormat::SAM[use_m & 1];
    ReadWriterSupplier *rws = ReadWriterSupplier::create(format, dws, genome, transcriptome, (const GTFReader *)gtf);
    ReadWriter *w = rws->getWriter();
    double t0 = now_s();
    for (unsigned i = 0; i < r0->n; i++) {
        Read read0, read1;
        make_read(&read0, r0, i, read_group);
        if (!r1) {
            if (a0[i].skip) continue;
            w->writeRead(&read0, (AlignmentResult)a0[i].status, a0[i].mapq, a0[i].location, (Direction)a0[i].direction, a0[i].is_transcriptome != 0,
                         a0[i].tlocation);
        } else {
            if (a0[i].skip || a1[i].skip) continue;
            make_read(&read1, r1, i, read_group);
            PairedAlignmentResult res;
            memset(&res, 0, sizeof(res));
            const snapb200_sam_alignment *a[2] = {&a0[i], &a1[i]};
            for (int e = 0; e < 2; e++) {
                res.status[e] = (AlignmentResult)a[e]->status;
                res.location[e] = a[e]->location;
                res.direction[e] = (Direction)a[e]->direction;
                res.mapq[e] = a[e]->mapq;
                res.isTranscriptome[e] = a[e]->is_transcriptome != 0;
                res.tlocation[e] = a[e]->tlocation;
            }
            w->writePair(&read0, &read1, &res);
        }
    }
    w->close();
    g_last_seconds = now_s() - t0;
    delete w;
    rws->close();
    delete rws;
    return 0;
}

int ref_sam_batch(void *h, const snapb200_sam_reads *r0, const snapb200_sam_reads *r1, const snapb200_sam_alignment *a0,
                  const snapb200_sam_alignment *a1, int use_m, const char *read_group, const char *path)
{
    return ref_sam_batch_rna(h, NULL, NULL, r0, r1, a0, a1, use_m, read_group, path);
}

// LandauVishkinWithCigar::insertSpliceJunctions (SNAPLib/LandauVishkin.cpp:119-250) on a token list (count, operator, count, ...) as
// computeCigarString leaves it; the string goes to out.  Returns what the reference returns (the number of operators, -2: no space).
int ref_splice_cigar(void *gtf, const unsigned *tokens, unsigned n_tokens, const char *transcript_id, unsigned pos, char *out, int out_len)
{
    LandauVishkinWithCigar lv;
    std::vector<unsigned> t(tokens, tokens + n_tokens);
    return lv.insertSpliceJunctions((const GTFReader *)gtf, t, std::string(transcript_id), pos, out, out_len);
}

// ---- row f3 (next): AlignmentFilter as the paired run loop drives it -------------------------------------------------------
// No CUDA counterpart exists yet; these entry points pin the reference's behaviour (tests/golden/filter_cases.npz) so that the
// device version has an oracle from its first line.

// GTFReader as AlignerContext creates it (SNAPLib/AlignerContext.cpp:265-267); out_prefix names the files WriteReadCounts writes.
void *ref_gtf_load(const char *gtf_path, const char *out_prefix)
{
    GTFReader *g = new GTFReader(strdup(out_prefix));
    g->Load(std::string(gtf_path));
    return g;
}

// AlignerContext::finishIteration's GTF epilogue (SNAPLib/AlignerContext.cpp:126-127)
int ref_gtf_finish(void *gtf)
{
    GTFReader *g = (GTFReader *)gtf;
    g->AnalyzeReadIntervals();
    g->WriteReadCounts();
    return 0;
}

struct ref_filter_result {  // the fields of PairedAlignmentResult that leave the loop (writePair / updateStats)
    unsigned location[2];
    unsigned tlocation[2];
    int score[2];
    int mapq[2];
    unsigned char status[2];
    unsigned char direction[2];
    unsigned char is_transcriptome[2];
    unsigned char aligned_as_pair;  // result.alignedAsPair after Filter (feeds the %Pairs column, PairedAligner.cpp:733-735)
    unsigned char pad;
};

// The part of PairedAlignerContext::runIterationThread between the aligner calls and writePair (SNAPLib/PairedAligner.cpp:
// 575-663): AlignmentFilter over the transcriptome multi-hits of both ends and the genome pair, Filter(), forceSpacing and the
// "cheese" MAPQ rule.  Alignments come in as the aligners (or the CUDA library) produced them; reads as the reader did.
// One thread, pairs in order: the filter updates shared GTF counters (SNAPLib/GTFReader.h) whose totals ref_gtf_finish writes.
int ref_filter_paired_batch(void *h_genome, void *h_transcriptome, void *gtf, const snapb200_sam_reads *r0, const snapb200_sam_reads *r1,
                            unsigned min_spacing, unsigned max_spacing, int force_spacing, unsigned conf_diff, unsigned max_dist,
                            unsigned max_hits_to_get, const int *n0, const unsigned *l0, const unsigned char *rc0, const int *sc0,
                            const int *n1, const unsigned *l1, const unsigned char *rc1, const int *sc1, const snapb200_paired_result *genome_res,
                            ref_filter_result *out)
{
    GenomeIndex *idx = (GenomeIndex *)h_genome, *tidx = (GenomeIndex *)h_transcriptome;
    GTFReader *g = (GTFReader *)gtf;
    // partialAligner, SNAPLib/PairedAligner.cpp:518-530 (explorePopularSeeds / stopOnFirstHit at their defaults, false)
    BaseAligner *partial = new BaseAligner(idx, 300, max_dist, MAX_READ_LENGTH, 12, 0.0, 2, NULL, NULL);
    for (unsigned i = 0; i < r0->n; i++) {
        Read read0, read1;
        make_read(&read0, r0, i, NULL);
        make_read(&read1, r1, i, NULL);
        PairedAlignmentResult result;
        memset(&result, 0, sizeof(result));
        for (int e = 0; e < 2; e++) {
            result.status[e] = (AlignmentResult)genome_res[i].status[e];
            result.location[e] = genome_res[i].location[e];
            result.direction[e] = (Direction)genome_res[i].direction[e];
            result.score[e] = genome_res[i].score[e];
            result.mapq[e] = genome_res[i].mapq[e];
            result.isTranscriptome[e] = false;
        }
        result.fromAlignTogether = genome_res[i].from_align_together != 0;
        result.alignedAsPair = genome_res[i].aligned_as_pair != 0;
        AlignmentFilter filter(&read0, &read1, idx->getGenome(), tidx->getGenome(), g, min_spacing, max_spacing, conf_diff, max_dist,
                               idx->getSeedLength(), partial);
        const size_t base = (size_t)i * max_hits_to_get;
        for (int k = 0; k < n0[i]; k++) filter.AddAlignment(l0[base + k], rc0[base + k] ? RC : FORWARD, sc0[base + k], 0, true, false);
        for (int k = 0; k < n1[i]; k++) filter.AddAlignment(l1[base + k], rc1[base + k] ? RC : FORWARD, sc1[base + k], 0, true, true);
        filter.AddAlignment(result.location[0], result.direction[0], result.score[0], result.mapq[0], false, false);
        filter.AddAlignment(result.location[1], result.direction[1], result.score[1], result.mapq[1], false, true);
        filter.Filter(&result);
        if (force_spacing && isOneLocation(result.status[0]) != isOneLocation(result.status[1])) {
            result.status[0] = result.status[1] = NotFound;
            result.location[0] = result.location[1] = InvalidGenomeLocation;
        }
        if (result.score[0] + result.score[1] >= 5) {  // "cheese"
            if (result.mapq[0] < 50) result.mapq[0] /= 2;
            if (result.mapq[1] < 50) result.mapq[1] /= 2;
        }
        memset(&out[i], 0, sizeof(out[i]));
        for (int e = 0; e < 2; e++) {
            out[i].location[e] = result.location[e];
            out[i].tlocation[e] = result.isTranscriptome[e] ? result.tlocation[e] : 0;
            out[i].score[e] = result.score[e];
            out[i].mapq[e] = result.mapq[e];
            out[i].status[e] = (unsigned char)result.status[e];
            out[i].direction[e] = (unsigned char)result.direction[e];
            out[i].is_transcriptome[e] = result.isTranscriptome[e] ? 1 : 0;
        }
        out[i].aligned_as_pair = result.alignedAsPair ? 1 : 0;
    }
    partial->~BaseAligner();
    return 0;
}

// How many intervals the filter has recorded so far (two per IntrachromosomalPair / ...Splice call): [intra pairs, intra splices,
// inter pairs, inter splices].  For sizing the host share of the replay (scripts/filter_host_profile.py).
int ref_gtf_interval_counts(void *gtf, unsigned long long *out)
{
    GTFReader *g = (GTFReader *)gtf;
    out[0] = g->intrachromosomal_pairs.read_intervals.size();
    out[1] = g->intrachromosomal_splices.read_intervals.size();
    out[2] = g->interchromosomal_pairs.read_intervals.size();
    out[3] = g->interchromosomal_splices.read_intervals.size();
    return 0;
}

// The annotation as the filter sees it, for checking a flat-table loader: one line per transcript
//   T <transcript_id> <chr> <gene_id> <start> <end> <n> then n x (type start end)     (GTFTranscript::exons order)
// and per gene   G <gene_id> <chr> <start> <end>.   Written to `path`.
int ref_gtf_export(void *gtf, const char *path)
{
    GTFReader *g = (GTFReader *)gtf;
    FILE *f = fopen(path, "w");
    if (!f) return -1;
    for (transcript_map::iterator it = g->transcripts.begin(); it != g->transcripts.end(); ++it) {
        GTFTranscript &t = it->second;
        fprintf(f, "T\t%s\t%s\t%s\t%u\t%u\t%u", it->first.c_str(), t.chr.c_str(), t.gene_id.c_str(), t.start, t.end, (unsigned)t.exons.size());
        for (feature_list::iterator e = t.exons.begin(); e != t.exons.end(); ++e) fprintf(f, "\t%u\t%u\t%u", (*e)->type, (*e)->start, (*e)->end);
        fprintf(f, "\n");
    }
    for (gene_map::iterator it = g->genes.begin(); it != g->genes.end(); ++it)
        fprintf(f, "G\t%s\t%s\t%u\t%u\n", it->first.c_str(), it->second.chr.c_str(), it->second.start, it->second.end);
    fclose(f);
    return 0;
}

// What AlignmentFilter::AddAlignment + HashAlignment leave in the two maps of one pair (AlignmentFilter.cpp:113-214), in map
// (string-key) order: for end e (0 = the map the run loop fills from read 0, i.e. the member called mate1), up to `cap` records of
//   location, pos, pos_end, pos_original, score, direction, isTranscriptome  (7 unsigned each) and the key strings, NUL-separated.
int ref_filter_alignments(void *h_genome, void *h_transcriptome, void *gtf, const snapb200_sam_reads *r0, const snapb200_sam_reads *r1,
                          unsigned i, unsigned max_dist, unsigned max_hits_to_get, const int *n0, const unsigned *l0, const unsigned char *rc0,
                          const int *sc0, const int *n1, const unsigned *l1, const unsigned char *rc1, const int *sc1,
                          const snapb200_paired_result *genome_res, unsigned cap, unsigned *counts, unsigned *records, char *keys, unsigned keys_cap)
{
    GenomeIndex *idx = (GenomeIndex *)h_genome, *tidx = (GenomeIndex *)h_transcriptome;
    Read read0, read1;
    make_read(&read0, r0, i, NULL);
    make_read(&read1, r1, i, NULL);
    AlignmentFilter filter(&read0, &read1, idx->getGenome(), tidx->getGenome(), (GTFReader *)gtf, 50, 1000, 2, max_dist, idx->getSeedLength(), NULL);
    const size_t base = (size_t)i * max_hits_to_get;
    for (int k = 0; k < n0[i]; k++) filter.AddAlignment(l0[base + k], rc0[base + k] ? RC : FORWARD, sc0[base + k], 0, true, false);
    for (int k = 0; k < n1[i]; k++) filter.AddAlignment(l1[base + k], rc1[base + k] ? RC : FORWARD, sc1[base + k], 0, true, true);
    filter.AddAlignment(genome_res[i].location[0], (Direction)genome_res[i].direction[0], genome_res[i].score[0], genome_res[i].mapq[0], false, false);
    filter.AddAlignment(genome_res[i].location[1], (Direction)genome_res[i].direction[1], genome_res[i].score[1], genome_res[i].mapq[1], false, true);
    alignment_map *maps[2] = {&filter.mate1, &filter.mate0};  // read 0's alignments live in the member called mate1 (isMate0 = false)
    unsigned kp = 0;
    for (int e = 0; e < 2; e++) {
        unsigned c = 0;
        for (alignment_map::iterator it = maps[e]->begin(); it != maps[e]->end(); ++it, ++c) {
            if (c >= cap) return -2;
            Alignment &a = it->second;
            unsigned *r = records + ((size_t)e * cap + c) * 7;
            r[0] = a.location; r[1] = a.pos; r[2] = a.pos_end; r[3] = a.pos_original; r[4] = (unsigned)a.score; r[5] = (unsigned)a.direction;
            r[6] = a.isTranscriptome ? 1 : 0;
            size_t len = it->first.size() + 1;
            if (kp + len > keys_cap) return -3;
            memcpy(keys + kp, it->first.c_str(), len);
            kp += (unsigned)len;
        }
        counts[e] = c;
    }
    return 0;
}

// The host half of a device filter: the GTF statistics of a batch, replayed in input order from per-pair event records (layout of
// FltEvent, snap_rnaseq_b200/csrc/filterfmt.h) through the reference's own public GTFReader methods -- what the shim will do once
// the decision itself comes from the device.  transcript_ids / chr_names: the strings behind the indices.
struct ref_flt_event { int kind, unaligned, transcript[2], chr[2]; unsigned pos_original[2], pos[2], pos_end[2]; };
struct ref_flt_splice { unsigned pair; int kind, chr[2]; unsigned pos[2], pos_end[2]; };
// splice_off / splices: when given, the novel-splice records of every pair (FltSplice, [splice_off[i], splice_off[i+1])) are handed to
// GTFReader::IntrachromosomalSplice / InterchromosomalSplice instead of running AlignmentFilter::UnalignedRead for the flagged reads
int ref_filter_replay_events2(void *h_genome, void *h_transcriptome, void *gtf, const snapb200_sam_reads *r0, const snapb200_sam_reads *r1,
                              unsigned max_dist, const ref_flt_event *ev, const char *const *transcript_ids, const char *const *chr_names,
                              const unsigned long long *splice_off, const ref_flt_splice *splices);
int ref_filter_replay_events(void *h_genome, void *h_transcriptome, void *gtf, const snapb200_sam_reads *r0, const snapb200_sam_reads *r1,
                             unsigned max_dist, const ref_flt_event *ev, const char *const *transcript_ids, const char *const *chr_names)
{
    return ref_filter_replay_events2(h_genome, h_transcriptome, gtf, r0, r1, max_dist, ev, transcript_ids, chr_names, NULL, NULL);
}
int ref_filter_replay_events2(void *h_genome, void *h_transcriptome, void *gtf, const snapb200_sam_reads *r0, const snapb200_sam_reads *r1,
                              unsigned max_dist, const ref_flt_event *ev, const char *const *transcript_ids, const char *const *chr_names,
                              const unsigned long long *splice_off, const ref_flt_splice *splices)
{
    GenomeIndex *idx = (GenomeIndex *)h_genome, *tidx = (GenomeIndex *)h_transcriptome;
    GTFReader *g = (GTFReader *)gtf;
    BaseAligner *partial = new BaseAligner(idx, 300, max_dist, MAX_READ_LENGTH, 12, 0.0, 2, NULL, NULL);
    for (unsigned i = 0; i < r0->n; i++) {
        Read read0, read1;
        make_read(&read0, r0, i, NULL);
        make_read(&read1, r1, i, NULL);
        const ref_flt_event &e = ev[i];
        if (e.unaligned && splice_off == NULL) {
            AlignmentFilter filter(&read0, &read1, idx->getGenome(), tidx->getGenome(), g, 50, 1000, 2, max_dist, idx->getSeedLength(), partial);
            filter.UnalignedRead(e.unaligned == 1 ? &read0 : &read1, idx->getSeedLength());
        } else if (e.unaligned) {
            Read *rd = e.unaligned == 1 ? &read0 : &read1;
            std::string rid(rd->getId(), rd->getIdLength());
            for (unsigned long long q = splice_off[i]; q < splice_off[i + 1]; q++) {
                const ref_flt_splice &sp = splices[q];
                if (sp.kind == 2) g->IntrachromosomalSplice(chr_names[sp.chr[0]], sp.pos[0], sp.pos_end[0], chr_names[sp.chr[1]], sp.pos[1], sp.pos_end[1], rid);
                else g->InterchromosomalSplice(chr_names[sp.chr[0]], sp.pos[0], sp.pos_end[0], chr_names[sp.chr[1]], sp.pos[1], sp.pos_end[1], rid);
            }
        }
        std::string t0 = e.transcript[0] >= 0 ? transcript_ids[e.transcript[0]] : "", t1 = e.transcript[1] >= 0 ? transcript_ids[e.transcript[1]] : "";
        std::string id(read0.getId(), read0.getIdLength());
        if (e.kind == 1) {  // AlignmentFilter.cpp:536-541: the lengths are passed crossed, as there
            g->IncrementReadCount(t0, e.pos_original[0], e.pos[0], read1.getDataLength(), t1, e.pos_original[1], e.pos[1], read0.getDataLength());
        } else if (e.kind == 2) {
            g->IntrachromosomalPair(chr_names[e.chr[0]], e.pos[0], e.pos_end[0], chr_names[e.chr[1]], e.pos[1], e.pos_end[1], id);
        } else if (e.kind == 3) {
            g->InterchromosomalPair(chr_names[e.chr[0]], e.pos[0], e.pos_end[0], chr_names[e.chr[1]], e.pos[1], e.pos_end[1], id);
        }
    }
    partial->~BaseAligner();
    return 0;
}

} // extern "C"




