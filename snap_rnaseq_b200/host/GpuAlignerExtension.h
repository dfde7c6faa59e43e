// GpuAlignerExtension.h -- reference-side binding of the B200 alignment core.
//
// This is the shim a SNAP-RNA maintainer adds to the reference tree (it compiles against the reference's
// own headers and links libsnapb200.so).  It uses the reference's plugin seam,
// `class AlignerExtension` (SNAPLib/AlignerContext.h:132-163), whose runIterationThread() may take over the
// whole per-thread loop (SNAPLib/SingleAligner.cpp:150-153, SNAPLib/PairedAligner.cpp:431-434):
//
//   1. drain the thread's ReadSupplier / PairedReadSupplier into batches (reads are only valid until the
//      next getNextRead(), Read.h:133-148, so ids/bases/qualities are copied into pinned SoA buffers);
//   2. call the C ABI (include/snapb200.h) where the reference calls BaseAligner::AlignRead /
//      ChimericPairedEndAligner::align;
//   3. replay the UNCHANGED host post-processing per read, in input order, exactly as
//      SingleAligner.cpp:243-304 / PairedAligner.cpp:547-668 do: AlignmentFilter, contamination filter, the
//      "cheese" MAPQ rule, writeRead/writePair (which computes the CIGAR), updateStats.
//
// Command lines, options and SAM output stay those of `snap-rna single|paired`.  Failures of the library
// map onto the reference's error model: message on stderr + soft_exit(1) (SNAPLib/exit.h:26).
#pragma once

#include <map>
#include <set>
#include <string>
#include <vector>

#include <pthread.h>
#include <sys/time.h>

#include "AlignerContext.h"
#include "AlignmentFilter.h"
#include "BaseAligner.h"
#include "PairedAligner.h"
#include "SingleAligner.h"
#include "WGsim.h"
#include "exit.h"

#include "snapb200.h"

// The reference grants `friend class AlignerContext2` in AlignerContext, SingleAlignerContext and
// PairedAlignerContext (AlignerContext.h:100, SingleAligner.h:63, PairedAligner.h:72) but never defines it:
// that name is the sanctioned way for an extension to reach writeRead/writePair/updateStats and the paired
// options.
class AlignerContext2 {
public:
    static void writeRead(SingleAlignerContext *c, Read *r, AlignmentResult res, unsigned loc, Direction d, bool isT, unsigned tloc, int score, int mapq)
    { c->writeRead(r, res, loc, d, isT, tloc, score, mapq); }
    static void updateStats(SingleAlignerContext *c, Read *r, AlignmentResult res, unsigned loc, int score, int mapq, bool err)
    { c->updateStats(c->stats, r, res, loc, score, mapq, err); }
    static void writePair(PairedAlignerContext *c, Read *r0, Read *r1, PairedAlignmentResult *res) { c->writePair(r0, r1, res); }
    static void updateStats(PairedAlignerContext *c, Read *r0, Read *r1, PairedAlignmentResult *res)
    { c->updateStats((PairedAlignerStats *)c->stats, r0, r1, res); }
    static int minSpacing(PairedAlignerContext *c) { return c->minSpacing; }
    static int maxSpacing(PairedAlignerContext *c) { return c->maxSpacing; }
    static bool forceSpacing(PairedAlignerContext *c) { return c->forceSpacing; }
    static unsigned maxBigHits(PairedAlignerContext *c) { return c->intersectingAlignerMaxHits; }
    static unsigned maxCandidatePoolSize(PairedAlignerContext *c) { return c->maxCandidatePoolSize; }
    static bool ignoreMismatchedIDs(PairedAlignerContext *c) { return c->ignoreMismatchedIDs; }
};

// BaseAligner::CharacterizeSeeds served from the device.  AlignmentFilter holds a
// `BaseAligner *specialAligner` and calls CharacterizeSeeds on it for unaligned / suspicious reads
// (SNAPLib/AlignmentFilter.cpp:758, 968-971); the extension computes the seed maps of every read of a batch with
// one snapb200_characterize_batch call per mate and this subclass hands the filter the maps of the read it asks
// about.  Needs the one-word change `virtual` on BaseAligner::CharacterizeSeeds (SNAPLib/BaseAligner.h:88); the
// base class's own CPU implementation is never run.
class GpuSeedCharacterizer : public BaseAligner {
public:
    GpuSeedCharacterizer(GenomeIndex *index, unsigned maxHits, unsigned maxK, unsigned nSeeds, double coverage, unsigned extraDepth,
                         bool explorePopular)
        : BaseAligner(index, maxHits, maxK, MAX_READ_LENGTH, nSeeds, coverage, extraDepth, NULL, NULL)
    {
        setExplorePopularSeeds(explorePopular);  // SingleAligner.cpp:183, PairedAligner.cpp:529
        params_.max_hits = maxHits; params_.max_k = maxK; params_.max_read_size = MAX_READ_LENGTH; params_.num_seeds = nSeeds;
        params_.seed_coverage = coverage; params_.extra_search_depth = extraDepth; params_.explore_popular_seeds = explorePopular ? 1 : 0;
        params_.stop_on_first_hit = 0; params_.max_hits_to_get = 0;
        current_[0] = current_[1] = NULL;
        index_ = 0;
    }

    // one device call per mate for the whole batch; returns the library's status
    int compute(snapb200_index *idx, int mate, const snapb200_read_batch *reads)
    {
        Maps &m = maps_[mate];
        m.seg.assign((size_t)2 * reads->n + 1, 0);
        int rc = snapb200_characterize_batch(idx, &params_, reads, &m.seg[0], NULL, NULL, 0);
        if (rc != SNAPB200_OK) return rc;
        const size_t total = (size_t)m.seg[(size_t)2 * reads->n];
        m.loc.resize(total + 1); m.off.resize(total + 1);
        return snapb200_characterize_batch(idx, &params_, reads, &m.seg[0], &m.loc[0], &m.off[0], total);
    }

    // the filter is about to be run on read i of the batch, whose mates live at r0 / r1 (r0 may be NULL: single end)
    void select(unsigned i, const Read *r0, const Read *r1) { index_ = i; current_[0] = r0; current_[1] = r1; }

    virtual AlignmentResult CharacterizeSeeds(Read *inputRead, unsigned *genomeLocation, Direction *hitDirection, int *finalScore,
                                              int *mapq, unsigned searchRadius, unsigned searchLocation, Direction searchDirection,
                                              seed_map &map, seed_map &mapRC)
    {
        (void)mapq; (void)searchRadius; (void)searchLocation; (void)searchDirection;
        *genomeLocation = InvalidGenomeLocation;   // what the reference leaves in its out-parameters (BaseAligner.cpp:250-252)
        if (hitDirection != NULL) *hitDirection = FORWARD;
        if (finalScore != NULL) *finalScore = 0xffff;
        const int mate = (inputRead == current_[1] && current_[0] != NULL) ? 1 : 0;
        const Maps &m = maps_[mate];
        seed_map *out[2] = {&map, &mapRC};
        for (int d = 0; d < 2; d++) {
            const size_t s = (size_t)2 * index_ + d;
            seed_map::iterator at = out[d]->end();
            for (uint64_t q = m.seg[s]; q < m.seg[s + 1]; q++) {  // ascending (location, offset): always appended at the end
                if (at == out[d]->end() || at->first != m.loc[q]) at = out[d]->insert(out[d]->end(), seed_map::value_type(m.loc[q], std::set<unsigned>()));
                at->second.insert(at->second.end(), (unsigned)m.off[q]);
            }
        }
        return NotFound;
    }

private:
    struct Maps { std::vector<uint64_t> seg; std::vector<uint32_t> loc; std::vector<uint16_t> off; };
    Maps maps_[2];
    snapb200_single_params params_;
    const Read *current_[2];
    unsigned index_;
};

class GpuAlignerExtension : public AlignerExtension {
public:
    explicit GpuAlignerExtension(int device = 0, unsigned batchReads = 1u << 17)
        : device_(device), batch_(batchReads), genome_(NULL), transcriptome_(NULL), contamination_(NULL), owner_(true) {}

    virtual ~GpuAlignerExtension() {}  // indices stay resident for the life of the process (see open())

    // One copy per worker thread (ParallelTask); the HBM-resident indices are shared, read-only.
    virtual AlignerExtension *copy()
    {
        GpuAlignerExtension *c = new GpuAlignerExtension(device_, batch_);
        c->genome_ = genome_; c->transcriptome_ = transcriptome_; c->contamination_ = contamination_;
        c->genomeDir_ = genomeDir_; c->transcriptomeDir_ = transcriptomeDir_; c->contaminationDir_ = contaminationDir_;
        c->owner_ = false;
        return c;
    }

    // AlignerContext::initialize() has loaded the host-side GenomeIndex objects; mirror them into HBM once.
    void attach(const AlignerOptions *options)
    {
        open(options->indexDir, &genome_, &genomeDir_);
        open(options->transcriptomeDir, &transcriptome_, &transcriptomeDir_);
        if (options->contaminationDir != NULL) open(options->contaminationDir, &contamination_, &contaminationDir_);
    }

    // ---- single end: replaces the loop at SNAPLib/SingleAligner.cpp:243-304 ----
    virtual bool runIterationThread(ReadSupplier *supplier, AlignerContext *ctx)
    {
        if (ctx->index == NULL) return false;  // "-" index: I/O only, leave it to the reference
        SingleAlignerContext *sc = (SingleAlignerContext *)ctx;
        if (genome_ == NULL) attach(ctx->options);
        snapb200_single_params p = singleParams(ctx);
        // the aligner AlignmentFilter calls CharacterizeSeeds on (SingleAligner.cpp:168 passes g_aligner); served from the device
        GpuSeedCharacterizer *partial = new GpuSeedCharacterizer(ctx->index, ctx->maxHits, ctx->maxDist, ctx->numSeedsFromCommandLine,
                                                                 ctx->seedCoverage, ctx->extraSearchDepth,
                                                                 ctx->options->explorePopularSeeds);
        ReadStore store;
        std::vector<snapb200_single_result> tres, gres, cres;
        Read *read;
        bool more = true;
        while (more) {
            store.clear();
            std::vector<char> useful;
            while (store.size() < batch_ && (more = (NULL != (read = supplier->getNextRead())))) {
                ctx->stats->totalReads++;
                bool quality = read->qualityFilter(ctx->options->minPercentAbovePhred, ctx->options->minPhred, ctx->options->phredOffset);
                bool ok = !(read->getDataLength() < 50 || read->countOfNs() > ctx->maxDist || !quality);  // SingleAligner.cpp:247-254
                store.add(read);
                useful.push_back(ok);
            }
            if (store.size() == 0) break;
            snapb200_read_batch rb = store.batch();
            tres.resize(store.size()); gres.resize(store.size());
            check(snapb200_single_batch(transcriptome_, &p, &rb, &tres[0]));
            check(snapb200_single_batch(genome_, &p, &rb, &gres[0]));
            check(partial->compute(genome_, 0, &rb));
            bool needContam = false;
            std::vector<AlignmentResult> final(store.size(), NotFound);
            // replay, in input order
            for (unsigned i = 0; i < store.size(); i++) {
                Read r;
                store.get(i, &r, ctx->clipping);
                if (!useful[i]) {
                    if (ctx->readWriter != NULL && ctx->options->passFilter(&r, NotFound))
                        ctx->readWriter->writeRead(&r, NotFound, 0, InvalidGenomeLocation, false, false, 0);
                    continue;
                }
                ctx->stats->usefulReads++;
                unsigned location = InvalidGenomeLocation, tlocation = 0;
                Direction direction = FORWARD;
                int score = 0, mapq = 0;
                bool isTranscriptome = false;
                partial->select(i, NULL, &r);
                AlignmentFilter filter(NULL, &r, ctx->index->getGenome(), ctx->transcriptome->getGenome(), ctx->gtf, 0, 0,
                                       ctx->options->confDiff, ctx->options->maxDist.start, ctx->index->getSeedLength(), partial);
                filter.AddAlignment(tres[i].location, tres[i].direction, tres[i].score, tres[i].mapq, true, true);
                filter.AddAlignment(gres[i].location, gres[i].direction, gres[i].score, gres[i].mapq, false, true);
                AlignmentResult result = filter.FilterSingle(&location, &direction, &score, &mapq, &isTranscriptome, &tlocation);
                if (result == NotFound && contamination_ != NULL) {
                    snapb200_read_batch one = store.one(i);
                    snapb200_single_result c;
                    check(snapb200_single_batch(contamination_, &p, &one, &c));
                    if (c.status != NotFound) ctx->c_filter->AddAlignment(c.location, c.direction, c.score, c.mapq, false, false);
                }
                bool wasError = false;
                if (result != NotFound && ctx->computeError) wasError = wgsimReadMisaligned(&r, location, ctx->index, ctx->options->misalignThreshold);
                AlignerContext2::writeRead(sc, &r, result, location, direction, isTranscriptome, tlocation, score, mapq);
                AlignerContext2::updateStats(sc, &r, result, location, score, mapq, wasError);
            }
            (void)needContam;
        }
        delete partial;
        return true;
    }

    // ---- paired end: replaces the loop at SNAPLib/PairedAligner.cpp:547-668 ----
    virtual bool runIterationThread(PairedReadSupplier *supplier, AlignerContext *ctx)
    {
        if (ctx->index == NULL) return false;
        PairedAlignerContext *pc = (PairedAlignerContext *)ctx;
        if (genome_ == NULL) attach(ctx->options);
        snapb200_paired_params pp;
        pp.max_hits = ctx->maxHits; pp.max_k = ctx->maxDist; pp.max_read_size = MAX_READ_LENGTH;
        pp.num_seeds = ctx->numSeedsFromCommandLine; pp.seed_coverage = ctx->seedCoverage;
        pp.min_spacing = AlignerContext2::minSpacing(pc); pp.max_spacing = AlignerContext2::maxSpacing(pc);
        pp.force_spacing = AlignerContext2::forceSpacing(pc); pp.max_big_hits = AlignerContext2::maxBigHits(pc);
        pp.extra_search_depth = ctx->extraSearchDepth; pp.max_candidate_pool_size = AlignerContext2::maxCandidatePoolSize(pc);
        snapb200_single_params tp = singleParams(ctx);  // transcriptomeAligner, PairedAligner.cpp:512
        const unsigned maxHitsToGet = 1000;             // PairedAligner.cpp:584
        tp.max_hits_to_get = maxHitsToGet;
        // partialAligner, PairedAligner.cpp:518-527: maxHits 300, 12 seeds; its CharacterizeSeeds is served from the device
        GpuSeedCharacterizer *partial = new GpuSeedCharacterizer(ctx->index, 300, ctx->maxDist, 12, ctx->seedCoverage, ctx->extraSearchDepth,
                                                                 ctx->options->explorePopularSeeds);
        ReadStore s0, s1;
        std::vector<snapb200_paired_result> res, cres;
        std::vector<snapb200_single_result> t0, t1;
        std::vector<int32_t> n0, n1, sc0, sc1;
        std::vector<uint32_t> l0, l1;
        std::vector<uint8_t> rc0, rc1;
        Read *read0, *read1;
        bool more = true;
        double tDrain = 0, tAbi = 0, tReplay = 0, tMark, tCall[3] = {0, 0, 0};
        unsigned long nReads = 0;
        while (more) {
            s0.clear(); s1.clear();
            std::vector<char> skip;
            tMark = now();
            while (s0.size() < batch_ && (more = supplier->getNextReadPair(&read0, &read1))) {
                if (!AlignerContext2::ignoreMismatchedIDs(pc)) Read::checkIdMatch(read0, read1);
                ctx->stats->totalReads += 2;
                int maxDist = ctx->maxDist;
                bool useful0 = read0->getDataLength() >= 50 && (int)read0->countOfNs() <= maxDist;
                bool useful1 = read1->getDataLength() >= 50 && (int)read1->countOfNs() <= maxDist;
                bool quality0 = read0->qualityFilter(ctx->options->minPercentAbovePhred, ctx->options->minPhred, ctx->options->phredOffset);
                bool bad = (!useful0 && !useful1) || (!quality0 || !quality0);  // sic, PairedAligner.cpp:564
                if (!bad) ctx->stats->usefulReads += (useful0 && useful1) ? 2 : 1;
                s0.add(read0); s1.add(read1);
                skip.push_back(bad);
            }
            const unsigned n = s0.size();
            tDrain += now() - tMark;
            if (n == 0) break;
            nReads += 2ul * n;
            tMark = now();
            snapb200_read_batch b0 = s0.batch(), b1 = s1.batch();
            res.resize(n); t0.resize(n); t1.resize(n); n0.resize(n); n1.resize(n);
            l0.resize((size_t)n * maxHitsToGet); l1.resize((size_t)n * maxHitsToGet); rc0.resize((size_t)n * maxHitsToGet);
            rc1.resize((size_t)n * maxHitsToGet); sc0.resize((size_t)n * maxHitsToGet); sc1.resize((size_t)n * maxHitsToGet);
            double tc = now();
            check(snapb200_single_multihit_batch(transcriptome_, &tp, &b0, &t0[0], &n0[0], &l0[0], &rc0[0], &sc0[0]));
            check(snapb200_single_multihit_batch(transcriptome_, &tp, &b1, &t1[0], &n1[0], &l1[0], &rc1[0], &sc1[0]));
            tCall[0] += now() - tc; tc = now();
            check(snapb200_paired_batch(genome_, &pp, &b0, &b1, &res[0]));
            tCall[1] += now() - tc; tc = now();
            check(partial->compute(genome_, 0, &b0));
            check(partial->compute(genome_, 1, &b1));
            tCall[2] += now() - tc;
            tAbi += now() - tMark;
            tMark = now();
            for (unsigned i = 0; i < n; i++) {
                Read r0, r1;
                s0.get(i, &r0, ctx->clipping); s1.get(i, &r1, ctx->clipping);
                PairedAlignmentResult result;
                if (skip[i]) {
                    result.status[0] = result.status[1] = NotFound;
                    result.location[0] = result.location[1] = InvalidGenomeLocation;
                    AlignerContext2::writePair(pc, &r0, &r1, &result);
                    continue;
                }
                toReference(res[i], &result);
                partial->select(i, &r0, &r1);
                AlignmentFilter filter(&r0, &r1, ctx->index->getGenome(), ctx->transcriptome->getGenome(), ctx->gtf, pp.min_spacing,
                                       pp.max_spacing, ctx->options->confDiff, ctx->options->maxDist.start, ctx->index->getSeedLength(), partial);
                for (int k = 0; k < n0[i]; k++) filter.AddAlignment(l0[(size_t)i * maxHitsToGet + k], rc0[(size_t)i * maxHitsToGet + k], sc0[(size_t)i * maxHitsToGet + k], 0, true, false);
                for (int k = 0; k < n1[i]; k++) filter.AddAlignment(l1[(size_t)i * maxHitsToGet + k], rc1[(size_t)i * maxHitsToGet + k], sc1[(size_t)i * maxHitsToGet + k], 0, true, true);
                filter.AddAlignment(result.location[0], result.direction[0], result.score[0], result.mapq[0], false, false);
                filter.AddAlignment(result.location[1], result.direction[1], result.score[1], result.mapq[1], false, true);
                filter.Filter(&result);
                if (result.status[0] == NotFound && result.status[1] == NotFound && contamination_ != NULL) {
                    snapb200_read_batch c0 = s0.one(i), c1 = s1.one(i);
                    snapb200_paired_result c;
                    check(snapb200_paired_batch(contamination_, &pp, &c0, &c1, &c));
                    if (c.status[0] != NotFound && c.status[1] != NotFound) {
                        ctx->c_filter->AddAlignment(c.location[0], c.direction[0], c.score[0], c.mapq[0], false, false);
                        ctx->c_filter->AddAlignment(c.location[1], c.direction[1], c.score[1], c.mapq[1], false, true);
                    }
                }
                if (pp.force_spacing && isOneLocation(result.status[0]) != isOneLocation(result.status[1])) {
                    result.status[0] = result.status[1] = NotFound;
                    result.location[0] = result.location[1] = InvalidGenomeLocation;
                }
                if (result.score[0] + result.score[1] >= 5) {  // "cheese", PairedAligner.cpp:653-663
                    if (result.mapq[0] < 50) result.mapq[0] /= 2;
                    if (result.mapq[1] < 50) result.mapq[1] /= 2;
                }
                AlignerContext2::writePair(pc, &r0, &r1, &result);
                AlignerContext2::updateStats(pc, &r0, &r1, &result);
            }
            tReplay += now() - tMark;
        }
        report("paired", tDrain, tAbi, tReplay, nReads);
        if (getenv("SNAPB200_SHIM_TIMING") != NULL)
            fprintf(stderr, "[snapb200 shim]   C ABI split: transcriptome multi-hit x2 %.2f s, paired %.2f s, CharacterizeSeeds x2 %.2f s, host buffers %.2f s\n",
                    tCall[0], tCall[1], tCall[2], tAbi - tCall[0] - tCall[1] - tCall[2]);
        snapb200_stats st;
        if (snapb200_stats_get(genome_, &st) == SNAPB200_OK) ctx->stats->lvCalls = st.n_locations_scored;
        delete partial;
        return true;
    }

private:
    // Owns copies of reads (id, unclipped bases, qualities) so they outlive the supplier's buffers, and
    // exposes the clipped reads as a snapb200_read_batch.
    struct ReadStore {
        std::vector<char> ids, bases, quals;          // unclipped, back to back
        std::vector<unsigned> idOff, off;             // n+1
        std::vector<uint8_t> cbases, cquals;          // clipped, what the aligner sees
        std::vector<uint32_t> coff;
        std::vector<const char *> readGroups;         // Read::getReadGroup(): owned by the reader context, outlives the batch
        ReadStore() { clear(); }
        void clear() { ids.clear(); bases.clear(); quals.clear(); cbases.clear(); cquals.clear(); readGroups.clear(); idOff.assign(1, 0); off.assign(1, 0); coff.assign(1, 0); }
        unsigned size() const { return (unsigned)off.size() - 1; }
        void add(Read *r)
        {
            ids.insert(ids.end(), r->getId(), r->getId() + r->getIdLength());
            idOff.push_back((unsigned)ids.size());
            bases.insert(bases.end(), r->getUnclippedData(), r->getUnclippedData() + r->getUnclippedLength());
            quals.insert(quals.end(), r->getUnclippedQuality(), r->getUnclippedQuality() + r->getUnclippedLength());
            off.push_back((unsigned)bases.size());
            cbases.insert(cbases.end(), (const uint8_t *)r->getData(), (const uint8_t *)r->getData() + r->getDataLength());
            cquals.insert(cquals.end(), (const uint8_t *)r->getQuality(), (const uint8_t *)r->getQuality() + r->getDataLength());
            coff.push_back((uint32_t)cbases.size());
            readGroups.push_back(r->getReadGroup());
        }
        void get(unsigned i, Read *r, ReadClippingType clipping)
        {
            r->init(&ids[idOff[i]], idOff[i + 1] - idOff[i], &bases[off[i]], &quals[off[i]], off[i + 1] - off[i]);
            r->clip(clipping);
            r->setReadGroup(readGroups[i]);
        }
        snapb200_read_batch batch() const
        {
            snapb200_read_batch b = {size(), &coff[0], cbases.empty() ? NULL : &cbases[0], cquals.empty() ? NULL : &cquals[0]};
            return b;
        }
        uint32_t oneOff[2];
        snapb200_read_batch one(unsigned i)
        {
            oneOff[0] = 0; oneOff[1] = coff[i + 1] - coff[i];
            snapb200_read_batch b = {1, oneOff, &cbases[coff[i]], &cquals[coff[i]]};
            return b;
        }
    };

    static void toReference(const snapb200_paired_result &g, PairedAlignmentResult *r)
    {
        for (int e = 0; e < 2; e++) {
            r->status[e] = (AlignmentResult)g.status[e]; r->location[e] = g.location[e]; r->direction[e] = g.direction[e];
            r->score[e] = g.score[e]; r->mapq[e] = g.mapq[e]; r->isTranscriptome[e] = false; r->tlocation[e] = 0;
        }
        r->fromAlignTogether = g.from_align_together != 0;
        r->alignedAsPair = g.aligned_as_pair != 0;
        r->nanosInAlignTogether = 0; r->nLVCalls = g.n_lv_calls; r->nSmallHits = 0;
    }

    snapb200_single_params singleParams(AlignerContext *ctx) const
    {
        snapb200_single_params p;
        p.max_hits = ctx->maxHits; p.max_k = ctx->maxDist; p.max_read_size = MAX_READ_LENGTH; p.num_seeds = ctx->numSeedsFromCommandLine;
        p.seed_coverage = ctx->seedCoverage; p.extra_search_depth = ctx->extraSearchDepth;
        p.explore_popular_seeds = ctx->options->explorePopularSeeds; p.stop_on_first_hit = ctx->options->stopOnFirstHit; p.max_hits_to_get = 0;
        return p;
    }

    // Process-wide cache of HBM-resident indices, like the reference's own index cache (AlignerContext.cpp:42-47):
    // every worker thread's copy of the extension ends up with the same handles.
    void open(const char *dir, snapb200_index **slot, std::string *name)
    {
        if (*slot != NULL && *name == dir) return;
        static pthread_mutex_t lock = PTHREAD_MUTEX_INITIALIZER;
        static std::map<std::string, snapb200_index *> cache;
        pthread_mutex_lock(&lock);
        std::map<std::string, snapb200_index *>::iterator it = cache.find(dir);
        if (it == cache.end()) {
            snapb200_index *h = NULL;
            check(snapb200_index_open(dir, device_, &h));
            it = cache.insert(std::make_pair(std::string(dir), h)).first;
        }
        *slot = it->second;
        *name = dir;
        pthread_mutex_unlock(&lock);
    }

    // SNAPB200_SHIM_TIMING=1: per worker thread, seconds spent draining the supplier, inside the C ABI and replaying the
    // host post-processing (printed to stderr when the thread finishes)
    static double now()
    {
        struct timeval tv;
        gettimeofday(&tv, NULL);
        return tv.tv_sec + tv.tv_usec * 1e-6;
    }
    static void report(const char *what, double drain, double abi, double replay, unsigned long reads)
    {
        if (getenv("SNAPB200_SHIM_TIMING") != NULL)
            fprintf(stderr, "[snapb200 shim] %s thread: %lu reads, drain %.2f s, C ABI %.2f s, host replay %.2f s\n", what, reads, drain, abi, replay);
    }

    static void check(int rc)
    {
        if (rc != SNAPB200_OK) {
            fprintf(stderr, "snapb200: %s\n", snapb200_last_error());
            soft_exit(1);
        }
    }

    int device_;
    unsigned batch_;
    snapb200_index *genome_, *transcriptome_, *contamination_;
    std::string genomeDir_, transcriptomeDir_, contaminationDir_;
    bool owner_;
};
