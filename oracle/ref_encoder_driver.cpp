// C entry point around the reference's own feature encoder -- TEST INFRASTRUCTURE (oracle/), NOT A PRODUCT PATH.
//
// oracle/Makefile compiles /root/reference/c++/src/{AlleleSearcherLiteFiltered,Read,Reference,Trie,leftAlignCigars,utils}.cpp
// where they lie, unmodified, against oracle/boost_shim (a stand-in for the Boost.Python container types; Boost itself is
// not in this image) and links them with this file into oracle/_ref/libref_encoder.so.  The function below builds an
// AlleleSearcherLiteFiltered exactly as python/PileupDataTools.py does (constructor, c++/src/AlleleSearcherLiteFiltered.cpp:
// 337-434), sets the three public members the allele search would have filled in (assemblyStart, assemblyStop,
// supports[allele]) and calls computeFeaturesColoredSimple (:1031-1180).  oracle/gen_encoder_cpp_golden.py uses it to pin
// oracle/encoder_oracle.py on the inputs the reference's Python specification cannot cover (soft clips, indels at the
// window border, insertions into reads of varying quality).
#include "AlleleSearcherLiteFiltered.h"

#include <cstring>

extern "C" long ref_encode_allele(int n_reads, const char* bases, const unsigned char* quals, const long* read_off,
                                  const int* cig_op, const int* cig_len, const long* cig_off, const long* ref_start,
                                  const int* mapq, const int* orient, const int* pacbio, const int* hp,
                                  const char* reference, long window_start, long assembly_start, long assembly_stop,
                                  const long* support_ids, int n_support, int feature_length, int pacbio_flag,
                                  int include_hp, unsigned char* out, long out_capacity_rows) {
    try {
        p::list reads, names, qualities, cigars, starts, mapqs, orients, pacbios, hps;
        for (int r = 0; r < n_reads; ++r) {
            reads.append(std::string(bases + read_off[r], bases + read_off[r + 1]));
            names.append(std::string("read") + std::to_string(r));
            p::list q;
            for (long k = read_off[r]; k < read_off[r + 1]; ++k) q.append((int)quals[k]);
            qualities.append(q);
            p::list c;
            for (long k = cig_off[r]; k < cig_off[r + 1]; ++k) {
                p::list op;
                op.append(cig_op[k]);
                op.append(cig_len[k]);
                c.append(op);
            }
            cigars.append(c);
            starts.append(ref_start[r]);
            mapqs.append(mapq[r]);
            orients.append(orient[r]);
            pacbios.append(pacbio[r] != 0);
            hps.append(hp[r]);
        }
        AlleleSearcherLiteFiltered searcher(reads, names, qualities, cigars, starts, mapqs, orients, pacbios, hps,
                                            std::string(reference), (size_t)window_start, (size_t)assembly_start,
                                            (size_t)assembly_stop, false);
        searcher.assemblyStart = (size_t)assembly_start;
        searcher.assemblyStop = (size_t)assembly_stop;
        const std::string allele = "allele";
        if (n_support > 0) searcher.supports[allele] = std::vector<size_t>(support_ids, support_ids + n_support);
        np::ndarray a = searcher.computeFeaturesColoredSimple(allele, (size_t)feature_length, pacbio_flag != 0, include_hp != 0);
        const long rows = a.shape(0);
        if (rows > out_capacity_rows) return -2;
        std::memcpy(out, a.get_data(), (size_t)rows * a.shape(1) * a.shape(2));
        return rows;
    } catch (const std::exception& e) {
        std::cerr << "ref_encode_allele: " << e.what() << std::endl;
        return -1;
    }
}
