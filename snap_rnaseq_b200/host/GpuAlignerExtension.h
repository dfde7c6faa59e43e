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

#include <malloc.h>
#include <pthread.h>
#include <sys/time.h>

#include "AlignerContext.h"
#include "AlignmentFilter.h"
#include "BaseAligner.h"
#include "DataWriter.h"
#include "FileFormat.h"
#include "GzipDataWriter.h"
#include "PairedAligner.h"
#include "SingleAligner.h"
#include "WGsim.h"
#include "exit.h"

#include "snapb200.h"

// SAM / BAM output whose records the device has already formatted (snapb200_rna_batch_submit_sam).  snap-rna-b200's main installs one
// of these over each FileFormat::SAM[] and FileFormat::BAM[] entry before runAlignment: it is the reference's format for everything (format detection,
// header, sort keys, writeRead for single reads and for pairs the shim leaves to the host) except that, while the calling thread has
// armed the two lines of a pair, writeRead copies them instead of formatting -- so SimpleReadWriter::writePair
// (SNAPLib/ReadWriter.cpp:132-217) keeps doing what it does: both lines into one buffer, a fresh buffer and a second try when they
// do not fit, writer->advance with the sort locations.
class PreformattedSAMFormat : public FileFormat {
public:
    struct Pending { const char *line[2]; size_t len[2]; bool armed; };
    static Pending &pending() { static __thread Pending p; return p; }  // zero-initialised per thread: not armed
    static bool &installed() { static bool yes = false; return yes; }
    static void install()
    {
        if (installed()) return;
        for (int k = 0; k < 2; k++) {
            FileFormat::SAM[k] = new PreformattedSAMFormat(FileFormat::SAM[k], false);
            FileFormat::BAM[k] = new PreformattedSAMFormat(FileFormat::BAM[k], true);
        }
        installed() = true;
    }
    // the two lines of the next writePair on this thread, in the order writePair writes them
    static void arm(const char *first, size_t firstLen, const char *second, size_t secondLen)
    {
        Pending &p = pending();
        p.line[0] = first; p.len[0] = firstLen; p.line[1] = second; p.len[1] = secondLen; p.armed = true;
    }
    static void disarm() { pending().armed = false; }

    PreformattedSAMFormat(const FileFormat *inner_, bool bam_) : inner(inner_), bam(bam_) {}
    virtual bool isFormatOf(const char *filename) const { return inner->isFormatOf(filename); }
    virtual void getSortInfo(const Genome *genome, char *buffer, _int64 bytes, unsigned *o_location, unsigned *o_readBytes, int *o_refID = NULL, int *o_pos = NULL) const
    { inner->getSortInfo(genome, buffer, bytes, o_location, o_readBytes, o_refID, o_pos); }
    // SAMFormat::getWriterSupplier (SNAPLib/SAM.cpp:688-707) / BAMFormat::getWriterSupplier (SNAPLib/Bam.cpp:508-539: the gzip filter,
    // duplicate marking and the index for sorted output) with this object as the writers' format
    virtual ReadWriterSupplier *getWriterSupplier(AlignerOptions *options, const Genome *genome, const Genome *transcriptome, const GTFReader *gtf) const
    {
        DataWriterSupplier *dataSupplier;
        if (bam) {
            GzipWriterFilterSupplier *gzipSupplier = DataWriterSupplier::gzip(true, 0x10000, options->numThreads, options->bindToProcessors, options->sortOutput);
            if (options->sortOutput) {
                const size_t len = strlen(options->outputFileTemplate);
                char *tempFileName = (char *)malloc(5 + len);
                strcpy(tempFileName, options->outputFileTemplate);
                strcpy(tempFileName + len, ".tmp");
                DataWriter::FilterSupplier *filters = gzipSupplier;
                if (!options->noDuplicateMarking) filters = DataWriterSupplier::markDuplicates(genome)->compose(filters);
                if (!options->noIndex) {
                    char *indexFileName = (char *)malloc(5 + len);
                    strcpy(indexFileName, options->outputFileTemplate);
                    strcpy(indexFileName + len, ".bai");
                    filters = DataWriterSupplier::bamIndex(indexFileName, genome, gzipSupplier)->compose(filters);
                }
                dataSupplier = DataWriterSupplier::sorted(this, genome, tempFileName, options->sortMemory * (1ULL << 30), options->numThreads,
                                                          options->outputFileTemplate, filters);
            } else {
                dataSupplier = DataWriterSupplier::create(options->outputFileTemplate, gzipSupplier);
            }
        } else if (options->sortOutput) {
            const size_t len = strlen(options->outputFileTemplate);
            char *tempFileName = (char *)malloc(5 + len);
            strcpy(tempFileName, options->outputFileTemplate);
            strcpy(tempFileName + len, ".tmp");
            dataSupplier = DataWriterSupplier::sorted(this, genome, tempFileName, options->sortMemory * (1ULL << 30), options->numThreads, options->outputFileTemplate, NULL);
        } else {
            dataSupplier = DataWriterSupplier::create(options->outputFileTemplate);
        }
        return ReadWriterSupplier::create(this, dataSupplier, genome, transcriptome, gtf);
    }
    virtual bool writeHeader(const ReaderContext &context, char *header, size_t headerBufferSize, size_t *headerActualSize, bool sorted, int argc, const char **argv,
                             const char *version, const char *rgLine) const
    { return inner->writeHeader(context, header, headerBufferSize, headerActualSize, sorted, argc, argv, version, rgLine); }
    virtual bool writeRead(const Genome *genome, const Genome *transcriptome, const GTFReader *gtf, LandauVishkinWithCigar *lv, char *buffer, size_t bufferSpace,
                           size_t *spaceUsed, size_t qnameLen, Read *read, AlignmentResult result, int mapQuality, unsigned genomeLocation, Direction direction,
                           bool isTranscriptome = false, unsigned tlocation = 0, bool hasMate = false, bool firstInPair = false, Read *mate = NULL,
                           AlignmentResult mateResult = NotFound, unsigned mateLocation = 0, Direction mateDirection = FORWARD, bool mateIsTranscriptome = false,
                           unsigned mateTlocation = 0) const
    {
        const Pending &p = pending();
        if (p.armed && hasMate) {
            const int k = firstInPair ? 0 : 1;
            if (p.len[k] > bufferSpace) return false;  // as SAMFormat::writeRead when snprintf runs out of space (SAM.cpp:1137-1142)
            memcpy(buffer, p.line[k], p.len[k]);
            if (spaceUsed != NULL) *spaceUsed = p.len[k];
            return true;
        }
        return inner->writeRead(genome, transcriptome, gtf, lv, buffer, bufferSpace, spaceUsed, qnameLen, read, result, mapQuality, genomeLocation, direction,
                                isTranscriptome, tlocation, hasMate, firstInPair, mate, mateResult, mateLocation, mateDirection, mateIsTranscriptome, mateTlocation);
    }

private:
    const FileFormat *inner;
    const bool bam;
};

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

    // ReadIntervalMap::AddInterval (SNAPLib/GTFReader.cpp:287-305) for many intervals: the two ReadInterval objects of every
    // candidate are built by the calling thread, outside any lock; the map's lock is taken once to append them all, in order.
    // (Needs the friend declarations of INTEGRATION.md section 3 in ReadInterval, ReadIntervalMap and GTFReader.)
    struct SpliceBatch {
        std::vector<ReadInterval *> mates[2];  // [0] intrachromosomal, [1] interchromosomal; consecutive entries are a pair
        // glibc grows a thread's malloc arena one mprotect (a write lock on the process's address space) per few pages; sixteen
        // worker threads allocating ~750 bytes per record spend most of their time waiting for each other there.  Asking for a
        // large block and freeing it again (with trimming disabled, see pregrowSetup) leaves the arena's top chunk that large, so
        // the next ~30 k records are carved out of it without a system call.
        size_t sinceGrow;
        SpliceBatch() : sinceGrow(~(size_t)0 / 2) {}
        static void pregrowSetup()
        {
            static bool done = false;
            if (done) return;
            done = true;
            mallopt(M_MMAP_THRESHOLD, 32 << 20);   // so that the 28 MB request below comes from the arena, not from mmap
            mallopt(M_TRIM_THRESHOLD, 1 << 30);    // and stays with the arena when it is freed
        }
        void pregrow()
        {
            if (getenv("SNAPB200_NO_PREGROW") != NULL) return;
            void *p = malloc(28 << 20);
            if (p != NULL) { memset(p, 0, 28 << 20); free(p); }  // touched here, once, by this thread: the page faults too are paid in bulk
            sinceGrow = 0;
        }
        void add(bool intra, const std::string &chr0, unsigned start0, unsigned end0, const std::string &chr1, unsigned start1, unsigned end1,
                 const std::string &id)
        {
            if (sinceGrow > (24u << 20)) pregrow();
            sinceGrow += 800;
            ReadInterval *m0 = new ReadInterval(chr0, start0, end0, id, true);
            ReadInterval *m1 = new ReadInterval(chr1, start1, end1, id, true);
            m0->mate.insert(m1);
            m1->mate.insert(m0);
            std::vector<ReadInterval *> &v = mates[intra ? 0 : 1];
            v.push_back(m0); v.push_back(m1);
        }
        void flush(GTFReader *gtf)
        {
            for (int k = 0; k < 2; k++) {
                if (mates[k].empty()) continue;
                ReadIntervalMap &m = k == 0 ? gtf->intrachromosomal_splices : gtf->interchromosomal_splices;
                AcquireExclusiveLock(&m.mutex);
                for (size_t q = 0; q < mates[k].size(); q++)
                    m.read_intervals.push_back(Interval<ReadInterval *>(mates[k][q]->start, mates[k][q]->end, mates[k][q]));
                ReleaseExclusiveLock(&m.mutex);
                mates[k].clear();
            }
        }
    };
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
        memset(view_, 0, sizeof(view_));
    }

    const snapb200_single_params &params() const { return params_; }

    // the seed tuples of a whole batch, as snapb200_rna_batch_wait hands them out (borrowed: valid until the next submit)
    void borrow(int mate, const uint64_t *seg, const uint32_t *loc, const uint16_t *off)
    {
        view_[mate].seg = seg; view_[mate].loc = loc; view_[mate].off = off;
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
        const View &m = view_[mate];
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
    struct View { const uint64_t *seg; const uint32_t *loc; const uint16_t *off; };
    View view_[2];
    snapb200_single_params params_;
    const Read *current_[2];
    unsigned index_;
};

// AlignmentFilter::UnalignedRead (the novel-splice search over the seed maps of a read without alignments, SNAPLib/AlignmentFilter.cpp:
// 742-933) is protected; the extension calls it for the reads the device filter flags (snapb200_filter_event::unaligned).
class FilterAccess : public AlignmentFilter {
public:
    FilterAccess(Read *r0, Read *r1, const Genome *g, const Genome *t, GTFReader *gtf, unsigned minSpacing, unsigned maxSpacing, unsigned confDiff,
                 unsigned maxDist, unsigned seedLen, BaseAligner *special)
        : AlignmentFilter(r0, r1, g, t, gtf, minSpacing, maxSpacing, confDiff, maxDist, seedLen, special) {}
    void unaligned(Read *read, unsigned minDiff) { UnalignedRead(read, minDiff); }
};

class GpuAlignerExtension : public AlignerExtension {
public:
    // batchReads: pairs (or reads) per device batch.  Devices: every visible GPU (SNAPB200_DEVICES=n limits it); batches are dealt
    // round-robin over them from this one process (GTF counters are process-global, so one process drives all).
    explicit GpuAlignerExtension(unsigned batchReads = 1u << 15) : batch_(batchReads), owner_(true), deviceSam_(false), samFlags_(0)
    {
        if (const char *e = getenv("SNAPB200_SHIM_BATCH")) { int v = atoi(e); if (v >= 16) batch_ = (unsigned)v; }  // tests: many small batches
        AlignerContext2::SpliceBatch::pregrowSetup();  // main thread (copies are made from this object), before the workers allocate
    }

    virtual ~GpuAlignerExtension() {}  // indices stay resident for the life of the process (see deviceSet())

    // One copy per worker thread (ParallelTask); the HBM-resident indices are shared, read-only.
    virtual AlignerExtension *copy()
    {
        GpuAlignerExtension *c = new GpuAlignerExtension(batch_);
        c->owner_ = false;
        return c;
    }

    // Starts placing the indices (and, for `paired`, the annotation tables) in HBM on a background thread, so that it overlaps the
    // reference's own loading of its host-side indices and GTF (AlignerContext::initialize).  argv: what follows the sub-command --
    // <genome-idx> <transcriptome-idx> <gtf> ... [-ct <contamination-idx>] (SingleAligner.cpp:52-72, PairedAligner.cpp:290-310).
    // Optional: without it the first worker thread opens the handles.
    static double sinceProcessStart() { return now() - processStart(); }  // the first call fixes the origin (snap-rna-b200's main makes it)

    static void prefetch(int argc, const char **argv, bool paired)
    {
        processStart();
        if (argc < 3 || strcmp(argv[0], "-") == 0) return;
        Prefetch *p = new Prefetch();
        p->index = argv[0]; p->transcriptome = argv[1]; p->annotation = argv[2]; p->paired = paired;
        for (int i = 3; i + 1 < argc; i++) {
            if (strcmp(argv[i], ",") == 0) break;
            if (strcmp(argv[i], "-ct") == 0) { p->contamination = argv[i + 1]; p->haveContamination = true; }
        }
        pthread_t th;
        if (pthread_create(&th, NULL, prefetchMain, p) == 0) pthread_detach(th); else delete p;
    }

    // ---- single end: replaces the loop at SNAPLib/SingleAligner.cpp:243-304 ----
    virtual bool runIterationThread(ReadSupplier *supplier, AlignerContext *ctx)
    {
        if (ctx->index == NULL) return false;  // "-" index: I/O only, leave it to the reference
        SingleAlignerContext *sc = (SingleAlignerContext *)ctx;
        const std::vector<DeviceSet> &devs = deviceSets(ctx->options, false);
        snapb200_single_params p = singleParams(ctx);
        // the aligner AlignmentFilter would call CharacterizeSeeds on (SingleAligner.cpp:168 passes g_aligner); FilterSingle never does
        GpuSeedCharacterizer *partial = new GpuSeedCharacterizer(ctx->index, ctx->maxHits, ctx->maxDist, ctx->numSeedsFromCommandLine,
                                                                 ctx->seedCoverage, ctx->extraSearchDepth,
                                                                 ctx->options->explorePopularSeeds);
        ReadStore store;
        std::vector<snapb200_single_result> tres, gres, cres;
        std::vector<uint32_t> contamIdx;
        Read *read;
        bool more = true;
        while (more) {
            store.clear();
            while (store.size() < batch_ && (more = (NULL != (read = supplier->getNextRead())))) {
                ctx->stats->totalReads++;
                bool quality = read->qualityFilter(ctx->options->minPercentAbovePhred, ctx->options->minPhred, ctx->options->phredOffset);
                bool ok = !(read->getDataLength() < 50 || read->countOfNs() > ctx->maxDist || !quality);  // SingleAligner.cpp:247-254
                store.add(read, ok);
            }
            if (store.size() == 0) break;
            const DeviceSet &dev = devs[nextDevice(devs.size())];
            snapb200_read_batch rb = store.batch();
            tres.resize(rb.n + 1); gres.resize(rb.n + 1);
            check(snapb200_single_batch(dev.transcriptome, &p, &rb, &tres[0]));
            check(snapb200_single_batch(dev.genome, &p, &rb, &gres[0]));
            // pass 1: the filter's verdict per read (no side effects on shared state); reads still unaligned go to the contamination index
            const unsigned n = store.size();
            std::vector<AlignmentResult> status(n, NotFound);
            std::vector<unsigned> location(n, InvalidGenomeLocation), tlocation(n, 0);
            std::vector<Direction> direction(n, FORWARD);
            std::vector<int> score(n, 0), mapq(n, 0);
            std::vector<char> isT(n, 0);
            contamIdx.clear();
            for (unsigned i = 0; i < n; i++) {
                const int di = store.deviceIndex(i);
                if (di < 0) continue;
                Read r;
                store.get(i, &r, ctx->clipping);
                bool isTranscriptome = false;
                AlignmentFilter filter(NULL, &r, ctx->index->getGenome(), ctx->transcriptome->getGenome(), ctx->gtf, 0, 0,
                                       ctx->options->confDiff, ctx->options->maxDist.start, ctx->index->getSeedLength(), partial);
                filter.AddAlignment(tres[di].location, tres[di].direction, tres[di].score, tres[di].mapq, true, true);
                filter.AddAlignment(gres[di].location, gres[di].direction, gres[di].score, gres[di].mapq, false, true);
                status[i] = filter.FilterSingle(&location[i], &direction[i], &score[i], &mapq[i], &isTranscriptome, &tlocation[i]);
                isT[i] = isTranscriptome;
                if (status[i] == NotFound && dev.contamination != NULL) contamIdx.push_back(i);
            }
            if (!contamIdx.empty()) {  // SingleAligner.cpp:282-294, batched
                ReadStore sub;
                for (size_t q = 0; q < contamIdx.size(); q++) sub.addFrom(store, contamIdx[q]);
                snapb200_read_batch cb = sub.batch();
                cres.resize(cb.n);
                check(snapb200_single_batch(dev.contamination, &p, &cb, &cres[0]));
                for (size_t q = 0; q < contamIdx.size(); q++)
                    if (cres[q].status != NotFound) ctx->c_filter->AddAlignment(cres[q].location, cres[q].direction, cres[q].score, cres[q].mapq, false, false);
            }
            // pass 2: output and statistics, in input order
            for (unsigned i = 0; i < n; i++) {
                Read r;
                store.get(i, &r, ctx->clipping);
                if (store.deviceIndex(i) < 0) {
                    if (ctx->readWriter != NULL && ctx->options->passFilter(&r, NotFound))
                        ctx->readWriter->writeRead(&r, NotFound, 0, InvalidGenomeLocation, false, false, 0);
                    continue;
                }
                ctx->stats->usefulReads++;
                bool wasError = false;
                if (status[i] != NotFound && ctx->computeError) wasError = wgsimReadMisaligned(&r, location[i], ctx->index, ctx->options->misalignThreshold);
                AlignerContext2::writeRead(sc, &r, status[i], location[i], direction[i], isT[i] != 0, tlocation[i], score[i], mapq[i]);
                AlignerContext2::updateStats(sc, &r, status[i], location[i], score[i], mapq[i], wasError);
            }
        }
        delete partial;
        return true;
    }

    // ---- paired end: replaces the loop at SNAPLib/PairedAligner.cpp:547-668 ----
    // Two batch objects per thread: while batch k is on a device (transcriptome multi-hits, genome pair, seed tuples and the
    // AlignmentFilter decision, all resident in HBM), the thread drains the supplier into batch k+1 and replays batch k-1: the
    // GTF counters through GTFReader's public methods, writePair, updateStats -- in input order.
    virtual bool runIterationThread(PairedReadSupplier *supplier, AlignerContext *ctx)
    {
        if (ctx->index == NULL) return false;
        PairedAlignerContext *pc = (PairedAlignerContext *)ctx;
        const double tEnter = now();
        const std::vector<DeviceSet> &devs = deviceSets(ctx->options, true);
        const double tInit = now() - tEnter;
        threadEnters();
        snapb200_rna_params P;
        snapb200_paired_params &pp = P.paired;
        pp.max_hits = ctx->maxHits; pp.max_k = ctx->maxDist; pp.max_read_size = MAX_READ_LENGTH;
        pp.num_seeds = ctx->numSeedsFromCommandLine; pp.seed_coverage = ctx->seedCoverage;
        pp.min_spacing = AlignerContext2::minSpacing(pc); pp.max_spacing = AlignerContext2::maxSpacing(pc);
        pp.force_spacing = AlignerContext2::forceSpacing(pc); pp.max_big_hits = AlignerContext2::maxBigHits(pc);
        pp.extra_search_depth = ctx->extraSearchDepth; pp.max_candidate_pool_size = AlignerContext2::maxCandidatePoolSize(pc);
        P.transcriptome = singleParams(ctx);             // transcriptomeAligner, PairedAligner.cpp:512
        P.transcriptome.max_hits_to_get = 1000;          // PairedAligner.cpp:584
        // partialAligner, PairedAligner.cpp:518-527: maxHits 300, 12 seeds; its CharacterizeSeeds is served from the device
        GpuSeedCharacterizer *partial = new GpuSeedCharacterizer(ctx->index, 300, ctx->maxDist, 12, ctx->seedCoverage, ctx->extraSearchDepth,
                                                                 ctx->options->explorePopularSeeds);
        P.partial = partial->params();
        P.filter.max_spacing = pp.max_spacing; P.filter.conf_diff = ctx->options->confDiff; P.filter.max_dist = ctx->options->maxDist.start;
        P.filter.max_hits_to_get = P.transcriptome.max_hits_to_get;
        P.filter.force_spacing = 0;  // applied below, after the contamination step, where the run loop applies it (PairedAligner.cpp:633-651)
        // SAM text from the device when the output goes through PreformattedSAMFormat (SNAPB200_HOST_SAM=1: the reference's writer formats)
        const bool samFile = ctx->options->outputFileTemplate != NULL && FileFormat::SAM[0]->isFormatOf(ctx->options->outputFileTemplate);
        const bool bamFile = ctx->options->outputFileTemplate != NULL && !samFile && FileFormat::BAM[0]->isFormatOf(ctx->options->outputFileTemplate);
        deviceSam_ = PreformattedSAMFormat::installed() && (samFile || bamFile) && getenv("SNAPB200_HOST_SAM") == NULL;
        samFlags_ = (ctx->options->useM ? SNAPB200_SAM_USE_M : 0) | (bamFile ? SNAPB200_SAM_BAM_RECORDS : 0);
        PairBatch pb[2];
        Timing tm;
        AlignerContext2::SpliceBatch splices;  // the thread's novel-splice intervals between two appends
        int cur = 0;
        fill(pb[cur], supplier, ctx, pc, tm);
        if (pb[cur].n) submit(pb[cur], devs, P);
        while (pb[cur].n) {
            const int nxt = cur ^ 1;
            fill(pb[nxt], supplier, ctx, pc, tm);       // host work that overlaps the device work of pb[cur]
            if (pb[nxt].n) submit(pb[nxt], devs, P);
            replay(pb[cur], devs, P, partial, ctx, pc, tm, splices);
            cur = nxt;
        }
        const double tLoop = now() - tEnter - tInit, td = now();
        for (int k = 0; k < 2; k++) pb[k].destroy();
        tm.report();
        if (getenv("SNAPB200_SHIM_TIMING") != NULL)
            fprintf(stderr, "[snapb200 shim]   thread wall: entered at %.2f s after process start, opening the device handles (or waiting for the thread that "
                            "does) %.2f s, batch loop %.2f s, releasing the batch objects %.2f s, left at %.2f s\n", tEnter - processStart(), tInit, tLoop,
                    now() - td, now() - processStart());
        ctx->stats->lvCalls = threadLeaves(devs);
        delete partial;
        return true;
    }

private:
    struct Prefetch { std::string index, transcriptome, annotation, contamination; bool haveContamination, paired; Prefetch() : haveContamination(false), paired(false) {} };
    static void *prefetchMain(void *arg)
    {
        Prefetch *p = (Prefetch *)arg;
        deviceSets(p->index.c_str(), p->transcriptome.c_str(), p->haveContamination ? p->contamination.c_str() : NULL, p->annotation.c_str(), p->paired);
        delete p;
        return NULL;
    }

    // Owns copies of reads (id, unclipped bases, qualities) so they outlive the supplier's buffers, and exposes the clipped reads
    // of the pairs that go to the device as a snapb200_read_batch.
    struct ReadStore {
        std::vector<char> ids, bases, quals;          // unclipped, back to back
        std::vector<unsigned> idOff, off;             // n+1
        std::vector<uint8_t> cbases, cquals;          // clipped, what the aligner sees (device reads only)
        std::vector<uint32_t> coff;
        std::vector<int> devIdx;                      // index in the device batch, -1: not aligned (the run loops' early-outs)
        std::vector<uint16_t> dFront, dClip;          // Read::getFrontClippedLength / getDataLength of the device reads (SAM stage)
        std::vector<char> gIds, gBases, gQuals;       // the unclipped device reads gathered, only when some read of the batch is not one
        std::vector<unsigned> gIdOff, gOff;
        std::vector<const char *> readGroups;         // Read::getReadGroup(): owned by the reader context, outlives the batch
        ReadStore() { clear(); }
        void clear()
        {
            ids.clear(); bases.clear(); quals.clear(); cbases.clear(); cquals.clear(); readGroups.clear(); devIdx.clear(); dFront.clear(); dClip.clear();
            idOff.assign(1, 0); off.assign(1, 0); coff.assign(1, 0);
        }
        unsigned size() const { return (unsigned)off.size() - 1; }
        unsigned deviceSize() const { return (unsigned)coff.size() - 1; }
        int deviceIndex(unsigned i) const { return devIdx[i]; }
        void add(Read *r, bool toDevice)
        {
            ids.insert(ids.end(), r->getId(), r->getId() + r->getIdLength());
            idOff.push_back((unsigned)ids.size());
            bases.insert(bases.end(), r->getUnclippedData(), r->getUnclippedData() + r->getUnclippedLength());
            quals.insert(quals.end(), r->getUnclippedQuality(), r->getUnclippedQuality() + r->getUnclippedLength());
            off.push_back((unsigned)bases.size());
            readGroups.push_back(r->getReadGroup());
            if (toDevice) {
                devIdx.push_back((int)deviceSize());
                cbases.insert(cbases.end(), (const uint8_t *)r->getData(), (const uint8_t *)r->getData() + r->getDataLength());
                cquals.insert(cquals.end(), (const uint8_t *)r->getQuality(), (const uint8_t *)r->getQuality() + r->getDataLength());
                coff.push_back((uint32_t)cbases.size());
                dFront.push_back((uint16_t)r->getFrontClippedLength());
                dClip.push_back((uint16_t)r->getDataLength());
            } else {
                devIdx.push_back(-1);
            }
        }
        // the device reads as snapb200_sam_reads: unclipped bases and qualities, ids, clipping
        snapb200_sam_reads samBatch()
        {
            static const uint8_t none8 = 0;
            static const uint16_t none16 = 0;
            snapb200_sam_reads s;
            s.n = deviceSize();
            if (s.n == size()) {
                s.offsets = &off[0]; s.id_offsets = &idOff[0];
                s.bases = bases.empty() ? &none8 : (const uint8_t *)&bases[0]; s.quals = quals.empty() ? &none8 : (const uint8_t *)&quals[0];
                s.ids = ids.empty() ? &none8 : (const uint8_t *)&ids[0];
            } else {
                gIds.clear(); gBases.clear(); gQuals.clear(); gIdOff.assign(1, 0); gOff.assign(1, 0);
                for (unsigned i = 0; i < size(); i++) {
                    if (devIdx[i] < 0) continue;
                    gIds.insert(gIds.end(), ids.begin() + idOff[i], ids.begin() + idOff[i + 1]);
                    gBases.insert(gBases.end(), bases.begin() + off[i], bases.begin() + off[i + 1]);
                    gQuals.insert(gQuals.end(), quals.begin() + off[i], quals.begin() + off[i + 1]);
                    gIdOff.push_back((unsigned)gIds.size()); gOff.push_back((unsigned)gBases.size());
                }
                s.offsets = &gOff[0]; s.id_offsets = &gIdOff[0];
                s.bases = gBases.empty() ? &none8 : (const uint8_t *)&gBases[0]; s.quals = gQuals.empty() ? &none8 : (const uint8_t *)&gQuals[0];
                s.ids = gIds.empty() ? &none8 : (const uint8_t *)&gIds[0];
            }
            s.front_clip = dFront.empty() ? &none16 : &dFront[0]; s.clipped_len = dClip.empty() ? &none16 : &dClip[0];
            return s;
        }
        // the clipped read i of another store, as a device read of this one (contamination sub-batches)
        void addFrom(const ReadStore &o, unsigned i)
        {
            const int d = o.devIdx[i];
            idOff.push_back(0); off.push_back(0); readGroups.push_back(NULL);
            devIdx.push_back((int)deviceSize());
            cbases.insert(cbases.end(), o.cbases.begin() + o.coff[d], o.cbases.begin() + o.coff[d + 1]);
            cquals.insert(cquals.end(), o.cquals.begin() + o.coff[d], o.cquals.begin() + o.coff[d + 1]);
            coff.push_back((uint32_t)cbases.size());
        }
        void get(unsigned i, Read *r, ReadClippingType clipping)
        {
            r->init(&ids[idOff[i]], idOff[i + 1] - idOff[i], &bases[off[i]], &quals[off[i]], off[i + 1] - off[i]);
            r->clip(clipping);
            r->setReadGroup(readGroups[i]);
        }
        snapb200_read_batch batch() const
        {
            static const uint8_t none = 0;
            snapb200_read_batch b = {deviceSize(), &coff[0], cbases.empty() ? &none : &cbases[0], cquals.empty() ? &none : &cquals[0]};
            return b;
        }
    };

    // The HBM-resident handles of one device.  Process-wide cache, like the reference's own index cache (AlignerContext.cpp:42-47):
    // every worker thread ends up with the same handles.
    struct DeviceSet {
        int device;
        snapb200_index *genome, *transcriptome, *contamination;
        snapb200_annotation *annotation;
        std::vector<std::string> transcriptIds;  // behind snapb200_filter_event::transcript
        std::vector<std::string> chrNames;       // genome piece names, behind the chr indices of events and splice records
    };

    struct OpenJob { const char *indexDir, *transcriptomeDir, *contaminationDir; DeviceSet set; };
    static void *openDevice(void *arg)
    {
        OpenJob *j = (OpenJob *)arg;
        DeviceSet &s = j->set;
        const double t0 = now();
        check(snapb200_index_open(j->indexDir, s.device, &s.genome));
        const double t1 = now();
        check(snapb200_index_open(j->transcriptomeDir, s.device, &s.transcriptome));
        if (j->contaminationDir != NULL) check(snapb200_index_open(j->contaminationDir, s.device, &s.contamination));
        if (getenv("SNAPB200_SHIM_TIMING") != NULL)
            fprintf(stderr, "[snapb200 shim] device %d: genome index in HBM after %.2f s (CUDA context creation included), the other indices %.2f s; "
                            "started %.2f s after process start\n", s.device, t1 - t0, now() - t1, t0 - processStart());
        return NULL;
    }

    static const std::vector<DeviceSet> &deviceSets(const AlignerOptions *options, bool needAnnotation)
    {
        return deviceSets(options->indexDir, options->transcriptomeDir, options->contaminationDir, options->annotation, needAnnotation);
    }

    static const std::vector<DeviceSet> &deviceSets(const char *indexDir, const char *transcriptomeDir, const char *contaminationDir, const char *annotation,
                                                    bool needAnnotation)
    {
        static pthread_mutex_t lock = PTHREAD_MUTEX_INITIALIZER;
        static std::vector<DeviceSet> sets;
        static std::string key;
        pthread_mutex_lock(&lock);
        std::string k = std::string(indexDir) + "\n" + transcriptomeDir + "\n" + (contaminationDir ? contaminationDir : "") + "\n" + annotation;
        if (sets.empty() || k != key) {
            // (a process that chains runs over different indices keeps the earlier handles resident, as the reference keeps its globals)
            sets.clear();
            int n = snapb200_device_count();
            if (const char *e = getenv("SNAPB200_DEVICES")) { int v = atoi(e); if (v >= 1 && v < n) n = v; }
            if (n < 1) { fprintf(stderr, "snapb200: no CUDA device available (this build has no CPU path)\n"); soft_exit(1); }
            // one opener thread per device: context creation and the index upload of the devices overlap (each GPU has its own link)
            std::vector<OpenJob> jobs(n);
            std::vector<pthread_t> threads(n);
            for (int d = 0; d < n; d++) {
                jobs[d].indexDir = indexDir; jobs[d].transcriptomeDir = transcriptomeDir; jobs[d].contaminationDir = contaminationDir;
                jobs[d].set.device = d; jobs[d].set.genome = jobs[d].set.transcriptome = jobs[d].set.contamination = NULL; jobs[d].set.annotation = NULL;
                if (n == 1) openDevice(&jobs[d]);
                else if (pthread_create(&threads[d], NULL, openDevice, &jobs[d]) != 0) { fprintf(stderr, "snapb200: pthread_create failed\n"); soft_exit(1); }
            }
            for (int d = 0; d < n; d++) {
                if (n > 1) pthread_join(threads[d], NULL);
                sets.push_back(jobs[d].set);
            }
            key = k;
        }
        if (needAnnotation && sets[0].annotation == NULL)
            for (size_t d = 0; d < sets.size(); d++) {
                check(snapb200_annotation_open(sets[d].genome, sets[d].transcriptome, annotation, &sets[d].annotation));
                const uint32_t nt = snapb200_annotation_transcript_count(sets[d].annotation);
                for (uint32_t t = 0; t < nt; t++) sets[d].transcriptIds.push_back(snapb200_annotation_transcript_id(sets[d].annotation, t));
                for (uint32_t c = 0; snapb200_annotation_chromosome(sets[d].annotation, c) != NULL; c++)
                    sets[d].chrNames.push_back(snapb200_annotation_chromosome(sets[d].annotation, c));
            }
        pthread_mutex_unlock(&lock);
        return sets;
    }

    // batches are dealt round-robin over the devices, whichever thread they come from (g = batch % nGPU)
    static size_t nextDevice(size_t nDevices)
    {
        static volatile unsigned counter = 0;
        return (size_t)(__sync_fetch_and_add(&counter, 1u) % (unsigned)nDevices);
    }

    // lvCalls (the stats line's column) is g_aligner->getLocationsScored(): locations scored by the intersecting aligner and its
    // single-end fallback on the genome index (ChimericPairedEndAligner.h:59-61).  The device keeps that counter per index, so the
    // thread that leaves last reports what has not been reported yet and the others report 0: the threads' sum is the total.
    static volatile int &activeThreads() { static volatile int v = 0; return v; }
    static volatile long long &reportedLv() { static volatile long long v = 0; return v; }
    static void threadEnters() { __sync_fetch_and_add(&activeThreads(), 1); }
    static long long threadLeaves(const std::vector<DeviceSet> &devs)
    {
        static pthread_mutex_t lock = PTHREAD_MUTEX_INITIALIZER;
        pthread_mutex_lock(&lock);
        long long mine = 0;
        if (__sync_sub_and_fetch(&activeThreads(), 1) == 0) {
            std::vector<snapb200_index *> copies;
            for (size_t d = 0; d < devs.size(); d++) copies.push_back(devs[d].genome);
            snapb200_stats st;
            const long long total = snapb200_stats_sum(&copies[0], (uint32_t)copies.size(), &st) == SNAPB200_OK ? st.n_locations_scored : 0;
            mine = total - reportedLv();
            reportedLv() = total;
        }
        pthread_mutex_unlock(&lock);
        return mine;
    }

    struct Timing {
        double drain, wait, replay, filterHost, unaligned, gtf, write, deviceMs;
        unsigned long devicePairs;  // pairs whose SAM lines the device formatted
        unsigned long reads, batches, hostPairs;
        std::vector<unsigned long> perDevice;  // batches each device took
        Timing() : drain(0), wait(0), replay(0), filterHost(0), unaligned(0), gtf(0), write(0), deviceMs(0), devicePairs(0), reads(0), batches(0), hostPairs(0) {}
        void report() const
        {
            if (getenv("SNAPB200_SHIM_TIMING") == NULL) return;
            fprintf(stderr, "[snapb200 shim] paired thread: %lu reads in %lu batches, drain %.2f s, waiting for the device %.2f s (device busy %.2f s), "
                            "host replay %.2f s (UnalignedRead %.2f, GTF counters %.2f, writePair+stats %.2f with the SAM lines of %lu pairs formatted on the device, "
                            "reference filter for %lu overflow pairs %.2f)\n",
                    reads, batches, drain, wait, deviceMs * 1e-3, replay, unaligned, gtf, write, devicePairs, hostPairs, filterHost);
            fprintf(stderr, "[snapb200 shim]   batches per device:");
            for (size_t d = 0; d < perDevice.size(); d++) fprintf(stderr, " gpu%zu=%lu", d, perDevice[d]);
            fprintf(stderr, "\n");
        }
    };

    struct PairBatch {
        ReadStore s0, s1;
        unsigned n;                      // pairs drained (device pairs: s0.deviceSize())
        std::vector<snapb200_rna_batch *> objs;  // one per device, created on first use
        int dev;
        bool sam;                        // submitted with the SAM stage
        PairBatch() : n(0), dev(0), sam(false) {}
        void destroy() { for (size_t d = 0; d < objs.size(); d++) if (objs[d]) snapb200_rna_batch_destroy(objs[d]); objs.clear(); }
    };

    void fill(PairBatch &b, PairedReadSupplier *supplier, AlignerContext *ctx, PairedAlignerContext *pc, Timing &tm)
    {
        const double t0 = now();
        b.s0.clear(); b.s1.clear();
        Read *read0, *read1;
        while (b.s0.size() < batch_ && supplier->getNextReadPair(&read0, &read1)) {
            if (!AlignerContext2::ignoreMismatchedIDs(pc)) Read::checkIdMatch(read0, read1);
            ctx->stats->totalReads += 2;
            int maxDist = ctx->maxDist;
            bool useful0 = read0->getDataLength() >= 50 && (int)read0->countOfNs() <= maxDist;
            bool useful1 = read1->getDataLength() >= 50 && (int)read1->countOfNs() <= maxDist;
            bool quality0 = read0->qualityFilter(ctx->options->minPercentAbovePhred, ctx->options->minPhred, ctx->options->phredOffset);
            bool bad = (!useful0 && !useful1) || (!quality0 || !quality0);  // sic, PairedAligner.cpp:564
            if (!bad) ctx->stats->usefulReads += (useful0 && useful1) ? 2 : 1;
            b.s0.add(read0, !bad); b.s1.add(read1, !bad);
        }
        b.n = b.s0.size();
        tm.drain += now() - t0;
        tm.reads += 2ul * b.n;
    }

    void submit(PairBatch &b, const std::vector<DeviceSet> &devs, const snapb200_rna_params &P)
    {
        b.dev = (int)nextDevice(devs.size());
        if (b.objs.size() < devs.size()) b.objs.resize(devs.size(), NULL);
        if (b.objs[b.dev] == NULL) check(snapb200_rna_batch_create(devs[b.dev].annotation, devs[b.dev].genome, devs[b.dev].transcriptome, &b.objs[b.dev]));
        snapb200_read_batch r0 = b.s0.batch(), r1 = b.s1.batch();
        // one read group per batch (FASTQ input: ReaderContext::defaultReadGroup for every read); otherwise the host formats this batch
        const char *group = NULL;
        bool uniform = true, first = true;
        for (unsigned i = 0; i < b.n && uniform; i++)
            for (int e = 0; e < 2; e++) {
                const ReadStore &st = e ? b.s1 : b.s0;
                if (st.devIdx[i] < 0) continue;
                const char *g = st.readGroups[i];
                if (g == READ_GROUP_FROM_AUX) { uniform = false; break; }
                if (first) { group = g; first = false; }
                else if (g != group && !(g != NULL && group != NULL && strcmp(g, group) == 0)) { uniform = false; break; }
            }
        b.sam = deviceSam_ && uniform;
        if (b.sam) {
            snapb200_sam_reads q0 = b.s0.samBatch(), q1 = b.s1.samBatch();
            check(snapb200_rna_batch_submit_sam(b.objs[b.dev], &P, &r0, &r1, &q0, &q1, samFlags_, group));
        } else {
            check(snapb200_rna_batch_submit(b.objs[b.dev], &P, &r0, &r1));
        }
    }

    void replay(PairBatch &b, const std::vector<DeviceSet> &devs, const snapb200_rna_params &P, GpuSeedCharacterizer *partial, AlignerContext *ctx,
                PairedAlignerContext *pc, Timing &tm, AlignerContext2::SpliceBatch &splices)
    {
        const DeviceSet &dev = devs[b.dev];
        const snapb200_paired_params &pp = P.paired;
        double t0 = now();
        snapb200_rna_view v;
        check(snapb200_rna_batch_wait(b.objs[b.dev], &v));
        tm.wait += now() - t0;
        tm.deviceMs += v.device_ms;
        tm.batches++;
        if (tm.perDevice.size() <= (size_t)b.dev) tm.perDevice.resize(b.dev + 1, 0);
        tm.perDevice[b.dev]++;
        t0 = now();
        const bool fine = getenv("SNAPB200_SHIM_TIMING") != NULL;
        for (int e = 0; e < 2; e++) partial->borrow(e, v.seg_offsets[e], v.ch_locations[e], v.ch_seed_offsets[e]);
        const unsigned seedLen = ctx->index->getSeedLength();
        const Genome *genome = ctx->index->getGenome();
        std::vector<PairedAlignmentResult> results(b.n);
        std::vector<unsigned> contam;
        const std::vector<std::string> &chrName = dev.chrNames;
        // pass 1: the pair's result and its GTF counters, in input order
        for (unsigned i = 0; i < b.n; i++) {
            PairedAlignmentResult &result = results[i];
            const int di = b.s0.deviceIndex(i);
            if (di < 0) {
                result.status[0] = result.status[1] = NotFound;
                result.location[0] = result.location[1] = InvalidGenomeLocation;
                continue;
            }
            Read r0, r1;
            b.s0.get(i, &r0, ctx->clipping); b.s1.get(i, &r1, ctx->clipping);
            if (v.needs_host[di]) {  // more alignments / combinations than the device scratch holds: the reference's class decides
                const double th = now();
                toReference(v.genome_pairs[di], &result);
                partial->select((unsigned)di, &r0, &r1);
                AlignmentFilter filter(&r0, &r1, genome, ctx->transcriptome->getGenome(), ctx->gtf, pp.min_spacing, pp.max_spacing, ctx->options->confDiff,
                                       ctx->options->maxDist.start, seedLen, partial);
                for (uint32_t k = v.hit_offsets[0][di]; k < v.hit_offsets[0][di + 1]; k++) filter.AddAlignment(v.hit_locations[0][k], v.hit_rcs[0][k] ? RC : FORWARD, v.hit_scores[0][k], 0, true, false);
                for (uint32_t k = v.hit_offsets[1][di]; k < v.hit_offsets[1][di + 1]; k++) filter.AddAlignment(v.hit_locations[1][k], v.hit_rcs[1][k] ? RC : FORWARD, v.hit_scores[1][k], 0, true, true);
                filter.AddAlignment(result.location[0], result.direction[0], result.score[0], result.mapq[0], false, false);
                filter.AddAlignment(result.location[1], result.direction[1], result.score[1], result.mapq[1], false, true);
                filter.Filter(&result);
                tm.filterHost += now() - th;
                tm.hostPairs++;
            } else {
                const snapb200_filter_result &fr = v.results[di];
                const snapb200_filter_event &ev = v.events[di];
                for (int e = 0; e < 2; e++) {
                    result.status[e] = (AlignmentResult)fr.status[e]; result.location[e] = fr.location[e]; result.direction[e] = (Direction)fr.direction[e];
                    result.score[e] = fr.score[e]; result.mapq[e] = fr.mapq[e]; result.isTranscriptome[e] = fr.is_transcriptome[e] != 0;
                    result.tlocation[e] = fr.tlocation[e];
                }
                result.fromAlignTogether = false;
                result.alignedAsPair = fr.aligned_as_pair != 0;
                result.nanosInAlignTogether = 0; result.nLVCalls = v.genome_pairs[di].n_lv_calls; result.nSmallHits = 0;
                if (ev.unaligned) {  // AlignmentFilter.cpp:331-340: the novel-splice search of the read that has no alignment at all
                    const double tu = fine ? now() : 0;
                    Read *ur = ev.unaligned == 1 ? &r0 : &r1;
                    if (v.splice_overflow[di]) {  // more partial alignments than the device scratch holds: the reference's own search
                        partial->select((unsigned)di, &r0, &r1);
                        FilterAccess fa(&r0, &r1, genome, ctx->transcriptome->getGenome(), ctx->gtf, pp.min_spacing, pp.max_spacing, ctx->options->confDiff,
                                        ctx->options->maxDist.start, seedLen, partial);
                        fa.unaligned(ur, seedLen);
                    } else if (v.splice_offsets[di + 1] > v.splice_offsets[di]) {
                        const std::string id(ur->getId(), ur->getIdLength());
                        for (uint64_t q = v.splice_offsets[di]; q < v.splice_offsets[di + 1]; q++) {
                            const snapb200_splice &sp = v.splices[q];
                            splices.add(sp.kind == 2, chrName[sp.chr[0]], sp.pos[0], sp.pos_end[0], chrName[sp.chr[1]], sp.pos[1], sp.pos_end[1], id);
                        }
                    }
                    if (fine) tm.unaligned += now() - tu;
                }
                if (ev.kind) {
                    const double tg = fine ? now() : 0;
                    if (ev.kind == 1) {  // AlignmentFilter.cpp:536-541 (the lengths are passed crossed there)
                        ctx->gtf->IncrementReadCount(ev.transcript[0] >= 0 ? dev.transcriptIds[ev.transcript[0]] : std::string(), ev.pos_original[0], ev.pos[0],
                                                     r1.getDataLength(), ev.transcript[1] >= 0 ? dev.transcriptIds[ev.transcript[1]] : std::string(),
                                                     ev.pos_original[1], ev.pos[1], r0.getDataLength());
                    } else {
                        const std::string id(r0.getId(), r0.getIdLength());
                        if (ev.kind == 2) ctx->gtf->IntrachromosomalPair(chrName[ev.chr[0]], ev.pos[0], ev.pos_end[0], chrName[ev.chr[1]], ev.pos[1], ev.pos_end[1], id);
                        else ctx->gtf->InterchromosomalPair(chrName[ev.chr[0]], ev.pos[0], ev.pos_end[0], chrName[ev.chr[1]], ev.pos[1], ev.pos_end[1], id);
                    }
                    if (fine) tm.gtf += now() - tg;
                }
            }
            if (result.status[0] == NotFound && result.status[1] == NotFound && dev.contamination != NULL) contam.push_back(i);
        }
        splices.flush(ctx->gtf);
        if (!contam.empty()) {  // PairedAligner.cpp:633-646, one device call for all unaligned pairs of the batch
            ReadStore c0, c1;
            for (size_t q = 0; q < contam.size(); q++) { c0.addFrom(b.s0, contam[q]); c1.addFrom(b.s1, contam[q]); }
            snapb200_read_batch r0 = c0.batch(), r1 = c1.batch();
            std::vector<snapb200_paired_result> cres(contam.size());
            check(snapb200_paired_batch(dev.contamination, &pp, &r0, &r1, &cres[0]));
            for (size_t q = 0; q < contam.size(); q++) {
                const snapb200_paired_result &c = cres[q];
                if (c.status[0] != NotFound && c.status[1] != NotFound) {
                    ctx->c_filter->AddAlignment(c.location[0], c.direction[0], c.score[0], c.mapq[0], false, false);
                    ctx->c_filter->AddAlignment(c.location[1], c.direction[1], c.score[1], c.mapq[1], false, true);
                }
            }
        }
        // pass 2: forceSpacing, the MAPQ halving where the host filter ran, output and statistics
        const double tw = now();
        for (unsigned i = 0; i < b.n; i++) {
            Read r0, r1;
            b.s0.get(i, &r0, ctx->clipping); b.s1.get(i, &r1, ctx->clipping);
            PairedAlignmentResult &result = results[i];
            const int di = b.s0.deviceIndex(i);
            if (di < 0) {
                AlignerContext2::writePair(pc, &r0, &r1, &result);
                continue;
            }
            if (pp.force_spacing && isOneLocation(result.status[0]) != isOneLocation(result.status[1])) {
                result.status[0] = result.status[1] = NotFound;
                result.location[0] = result.location[1] = InvalidGenomeLocation;
            }
            if (v.needs_host[di] && result.score[0] + result.score[1] >= 5) {  // "cheese", PairedAligner.cpp:653-663 (the device applied it already)
                if (result.mapq[0] < 50) result.mapq[0] /= 2;
                if (result.mapq[1] < 50) result.mapq[1] /= 2;
            }
            if (b.sam && v.sam_line_offsets != NULL) {
                // the pair's two lines came back with the batch: writePair only places them (an empty range: the host formats)
                const uint64_t *lo = v.sam_line_offsets + 2 * (size_t)di;
                if (lo[1] > lo[0] && lo[2] > lo[1]) {
                    PreformattedSAMFormat::arm(v.sam_text + lo[0], (size_t)(lo[1] - lo[0]), v.sam_text + lo[1], (size_t)(lo[2] - lo[1]));
                    tm.devicePairs++;
                }
            }
            AlignerContext2::writePair(pc, &r0, &r1, &result);
            PreformattedSAMFormat::disarm();
            AlignerContext2::updateStats(pc, &r0, &r1, &result);
        }
        tm.write += now() - tw;
        tm.replay += now() - t0;
    }

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

    static double processStart() { static const double t = now(); return t; }

    static double now()
    {
        struct timeval tv;
        gettimeofday(&tv, NULL);
        return tv.tv_sec + tv.tv_usec * 1e-6;
    }

    static void check(int rc)
    {
        if (rc != SNAPB200_OK) {
            fprintf(stderr, "snapb200: %s\n", snapb200_last_error());
            soft_exit(1);
        }
    }

    unsigned batch_;
    bool owner_;
    bool deviceSam_;  // paired loop: the SAM lines / BAM records come back with the batch (PreformattedSAMFormat installed, output .sam or .bam)
    int samFlags_;    // SNAPB200_SAM_USE_M, SNAPB200_SAM_BAM_RECORDS
};
