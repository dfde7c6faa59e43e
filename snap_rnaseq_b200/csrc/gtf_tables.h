// gtf_tables.h -- the annotation as flat tables for the device AlignmentFilter (row f3; host code, used by
// snapb200_annotation_open in filter_api.inl).  Restates what GTFReader::Load / Parse / GTFGene::Process / GTFTranscript::Process leave behind
// (SNAPLib/GTFReader.cpp:646-713, 857-872, 972-1019, 1245-1361) for the parts the filter reads: per transcript its chromosome, gene,
// extent and the exon / intron list GenomicPosition walks; per gene its chromosome and extent.  The reference's behaviour is kept where
// it decides results, including what looks like accidents.
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <fstream>
#include <map>
#include <set>
#include <string>
#include <vector>

struct GtfFeatureRow { uint32_t type, start, end; };  // type: 1 exon, 2 intron (SNAPLib/GTFReader.h:49)

struct GtfTranscriptRow {
    std::string id, chr, gene_id;
    uint32_t start, end;
    std::vector<GtfFeatureRow> features;  // GTFTranscript::exons: exons with the introns between them; empty if never processed
};

struct GtfGeneRow { std::string id, chr; uint32_t start, end; };

struct GtfTables {
    std::vector<GtfTranscriptRow> transcripts;  // in transcript-id order (std::map order of the reference)
    std::vector<GtfGeneRow> genes;              // in gene-id order
};

namespace gtf_detail {
// strtok as GTFFeature::GTFFeature uses it: skip leading delimiters, cut at the next one (which is consumed)
struct Tok {
    std::string s;
    size_t p = 0;
    bool next(const char *delims, std::string *out)
    {
        while (p < s.size() && strchr(delims, s[p])) p++;
        if (p >= s.size()) return false;
        size_t e = p;
        while (e < s.size() && !strchr(delims, s[e])) e++;
        *out = s.substr(p, e - p);
        p = e < s.size() ? e + 1 : e;
        return true;
    }
};

struct Feature {
    std::string chr, key, gene_id, transcript_id;
    uint32_t start = 0, end = 0;
};

struct Transcript {
    std::string chr, gene_id;
    uint32_t start, end;
    std::vector<const Feature *> features;
    std::vector<GtfFeatureRow> exons;
};

struct Gene {
    std::string chr;
    uint32_t start, end;
    std::set<std::string> transcript_ids;
};

inline bool by_start(const Feature *a, const Feature *b) { return a->start < b->start; }
}  // namespace gtf_detail

// Returns false if the file cannot be read.
inline bool gtf_load_tables(const char *path, GtfTables *out)
{
    using namespace gtf_detail;
    std::ifstream in(path, std::ios::in);
    if (!in.is_open()) return false;
    std::map<std::string, Feature> features;
    std::map<std::string, Transcript> transcripts;
    std::map<std::string, Gene> genes;
    std::string line;
    std::getline(in, line, '\n');
    while (!in.eof()) {  // as the reference reads: a last line without a newline is not parsed
        if (!line.empty() && line[0] != '#') {
            Tok t;
            t.s = line;
            const char *tab = "'\t'";  // the reference's delimiter set: tab AND the single quote
            std::string chr, source, feature, s_start, s_end, score, strand, frame;
            const bool ok = t.next(tab, &chr) && t.next(tab, &source) && t.next(tab, &feature) && t.next(tab, &s_start) && t.next(tab, &s_end) &&
                            t.next(tab, &score) && t.next(tab, &strand) && t.next(tab, &frame);
            if (ok && feature == "exon") {
                std::map<std::string, std::string> attr;
                std::string k, v;
                while (t.next(" =", &k)) {
                    if (!t.next(";", &v)) v.clear();
                    v.erase(std::remove(v.begin(), v.end(), '"'), v.end());
                    attr.insert(std::make_pair(k, v));
                }
                Feature f;
                f.chr = chr;
                f.start = (uint32_t)atoi(s_start.c_str());
                f.end = (uint32_t)atoi(s_end.c_str());
                if (attr.count("gene_id")) f.gene_id = attr["gene_id"];
                else if (attr.count("Parent")) f.gene_id = attr["Parent"];
                else f.gene_id = "Unknown";
                f.transcript_id = attr.count("transcript_id") ? attr["transcript_id"] : f.gene_id;
                f.key = f.gene_id + chr + s_start + s_end;  // text concatenation, as in the reference
                const Feature *shared = &features.insert(std::make_pair(f.key, f)).first->second;  // the first feature with a key stays
                std::map<std::string, Transcript>::iterator tp = transcripts.find(f.transcript_id);
                if (tp == transcripts.end()) {
                    Transcript tr;
                    tr.chr = f.chr; tr.gene_id = f.gene_id; tr.start = f.start; tr.end = f.end;
                    tr.features.push_back(shared);
                    transcripts.insert(std::make_pair(f.transcript_id, tr));
                } else {
                    tp->second.features.push_back(shared);
                    tp->second.start = std::min(tp->second.start, f.start);
                    tp->second.end = std::max(tp->second.end, f.end);
                }
                std::map<std::string, Gene>::iterator gp = genes.find(f.gene_id);
                if (gp == genes.end()) {
                    Gene g;
                    g.chr = f.chr; g.start = f.start; g.end = f.end;
                    // the reference inserts the transcript id before copying the gene into its map, and GTFGene's copy constructor
                    // drops transcript_ids (GTFReader.cpp:822-826): the id of a gene's first line is NOT registered here
                    genes.insert(std::make_pair(f.gene_id, g));
                } else {
                    gp->second.transcript_ids.insert(f.transcript_id);
                    gp->second.start = std::min(gp->second.start, f.start);
                    gp->second.end = std::max(gp->second.end, f.end);
                }
            }
        }
        std::getline(in, line, '\n');
    }
    // GTFGene::Process -> GTFTranscript::Process for the registered transcripts, genes in id order
    for (std::map<std::string, Gene>::iterator g = genes.begin(); g != genes.end(); ++g) {
        for (std::set<std::string>::iterator id = g->second.transcript_ids.begin(); id != g->second.transcript_ids.end(); ++id) {
            Transcript &tr = transcripts.find(*id)->second;
            std::sort(tr.features.begin(), tr.features.end(), by_start);  // by start only; libstdc++'s order for equal starts, as there
            const Feature *prev = NULL;
            for (size_t k = 0; k < tr.features.size(); k++) {
                const Feature *cur = tr.features[k];
                if (prev) {
                    GtfFeatureRow intron = {2u, prev->end + 1, cur->start - 1};
                    tr.exons.push_back(intron);
                }
                GtfFeatureRow exon = {1u, cur->start, cur->end};
                tr.exons.push_back(exon);
                prev = cur;
            }
        }
    }
    out->transcripts.clear();
    out->genes.clear();
    for (std::map<std::string, Transcript>::iterator t = transcripts.begin(); t != transcripts.end(); ++t) {
        GtfTranscriptRow r;
        r.id = t->first; r.chr = t->second.chr; r.gene_id = t->second.gene_id; r.start = t->second.start; r.end = t->second.end;
        r.features = t->second.exons;
        out->transcripts.push_back(r);
    }
    for (std::map<std::string, Gene>::iterator g = genes.begin(); g != genes.end(); ++g) {
        GtfGeneRow r;
        r.id = g->first; r.chr = g->second.chr; r.start = g->second.start; r.end = g->second.end;
        out->genes.push_back(r);
    }
    return true;
}
