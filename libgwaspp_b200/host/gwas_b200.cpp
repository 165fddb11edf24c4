// Minimal harness in the shape of the reference's GWAS executable (src/test/gwas_basic.cpp:126-244) for
// the association flags only: loads a transposed-PLINK pair (TPED/TFAM; the TPED may be .gz, or a SNP-major
// .bed), then runs one test-class function through compute(). The genotype file goes to the device as text
// and is parsed there (csrc/ingest.cu: alleles 1/2/3/4 -> A/C/G/T and pair collapse as in
// genetics/individual/tped_genotype_file.cpp:127-190); the phenotype file is read here, column 6 with
// '1' = case, '0' = control as in tfam_annotation_file.cpp:68-77.
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "test_functions.h"

using namespace libgwaspp::genetics;
using namespace libgwaspp::algorithms;

static void usage() {
    std::cerr << "usage: gwas_b200 -g X.tped -p X.tfam [-o out] [--device N] [--devices N] [--comp-level 5] "
                 "(--test-inline-maf | --select-cc-maf | --inline-cc-maf | --dist-perform | --test-boost-epi | "
                 "--contin-debug | --contin-perform | --contin-cc-perform | --epi-debug | --epi-perform)\n";
}

static int run_test(GeneticData &gd, const std::set<int> &cases, const std::set<int> &controls, const std::string &test, const std::string &outfile) {
    gd.setCaseControlSet(cases, controls);
    std::cout << "Setting " << cases.size() << " cases." << std::endl;
    std::cout << "Setting " << controls.size() << " controls." << std::endl;

    std::ofstream of;
    std::ostream *out = &std::cout;
    if (!outfile.empty()) { of.open(outfile.c_str()); if (!of.is_open()) { std::cerr << "cannot open " << outfile << std::endl; return 1; } out = &of; }

    if (test == "dump-api") {
        // every per-call virtual of the GenoTable interface, one line each (used by tests/test_gpu_host_layer.py)
        DeviceGenoTable &gt = *gd.getGenotypeTable();
        CaseControlSet &ccs = *gd.getCaseControlSet();
        const int M = gt.row_size(), N = gt.column_size();
        for (int r = 0; r < M; r += 7)
            for (int c = 0; c < N; c += 11) *out << "call " << r << " " << c << " " << gt.getCallAt((uint)r, (uint)c) << "\n";
        auto ft = [&](const char *tag, int r, const frequency_table &a) { *out << tag << " " << r << " " << a.aa << " " << a.ab << " " << a.bb << " " << a.xx << "\n"; };
        for (int r = 0; r < M; ++r) { GenotypeDistribution d; gt.getGenotypeDistribution((uint)r, d); ft("whole", r, *d.getDistribution()); }
        for (int r = 0; r < M; ++r) { CaseControlGenotypeDistribution d; gt.getCaseControlGenotypeDistribution((uint)r, ccs, d); ft("mask_ca", r, *d.getCaseDistribution()); ft("mask_co", r, *d.getControlDistribution()); }
        gt.selectCaseControl(ccs);
        marginal_information *mar = NULL; int nm = 0;
        computeMargins(gt, N, mar, nm);
        for (int r = 0; r < M; ++r) {
            CaseControlGenotypeDistribution d; gt.getCaseControlGenotypeDistribution((uint)r, d); ft("sel_ca", r, *d.getCaseDistribution()); ft("sel_co", r, *d.getControlDistribution());
            marginal_information m; CaseControlGenotypeDistribution e; gt.getCaseControlGenotypeDistribution((uint)r, e, m);
            ft("mar_ca", r, m.cases); ft("mar_co", r, m.controls);
            if (memcmp(&m, &mar[r], sizeof m) != 0) *out << "MARGIN_MISMATCH " << r << "\n";
        }
        auto tab = [&](const char *tag, int i, int j, const CONTIN_TABLE_T &t) { *out << tag << " " << i << " " << j; for (int q = 0; q < 16; ++q) *out << " " << t.contin[q]; *out << "\n"; };
        for (int i = 0; i < M; i += 5)
            for (int j = i + 1; j < M; j += 9) {
                ContingencyTable ct; gt.getContingencyTable((uint)i, (uint)j, ct); tab("t0", i, j, *ct.getContingencyTable());
                CaseControlContingencyTable a, b, c;
                gt.getCaseControlContingencyTable((uint)i, (uint)j, ccs, a); tab("t1_ca", i, j, *a.getCaseContingencyTable()); tab("t1_co", i, j, *a.getControlContingencyTable());
                gt.getCaseControlContingencyTable((uint)i, (uint)j, b); tab("t2_ca", i, j, *b.getCaseContingencyTable()); tab("t2_co", i, j, *b.getControlContingencyTable());
                gt.getCaseControlContingencyTable((uint)i, (uint)j, mar[i], mar[j], c); tab("t3_ca", i, j, *c.getCaseContingencyTable()); tab("t3_co", i, j, *c.getControlContingencyTable());
            }
        delete[] mar;
    }
    else if (test == "test-inline-maf") compute(inline_maf_print, &gd, out);
    else if (test == "select-cc-maf") compute(select_cc_maf, &gd, out);
    else if (test == "inline-cc-maf") compute(inline_cc_maf, &gd, out);
    else if (test == "dist-perform") compute(genotype_dist_performance, &gd, out);
    else if (test == "test-boost-epi") compute(computeBoost, &gd, out);
    else if (test == "contin-debug") compute(ContingencyDebug, &gd, out);
    else if (test == "contin-perform") compute(ContingencyPerformance, &gd, out);
    else if (test == "contin-cc-perform") compute(ContingencyCCPerformance, &gd, out);
    else if (test == "epi-debug") compute(EpistasisDebug, &gd, out);
    else if (test == "epi-perform") compute(EpistasisPerformance, &gd, out);
    else { usage(); return 1; }
    std::cout << "DONE" << std::endl;
    return 0;
}

int main(int argc, char **argv) {
    std::string geno, pheno, outfile, test;
    int device = 0, devices = 1, comp_level = 5;
    bool host_parse = false;
    for (int a = 1; a < argc; ++a) {
        std::string s = argv[a];
        if ((s == "-g" || s == "--geno") && a + 1 < argc) geno = argv[++a];
        else if ((s == "-p" || s == "--pheno") && a + 1 < argc) pheno = argv[++a];
        else if ((s == "-o" || s == "--output") && a + 1 < argc) outfile = argv[++a];
        else if (s == "--device" && a + 1 < argc) device = atoi(argv[++a]);
        else if (s == "--devices" && a + 1 < argc) devices = atoi(argv[++a]);
        else if (s == "--comp-level" && a + 1 < argc) comp_level = atoi(argv[++a]);
        else if (s == "--tplink") {}
        else if (s == "--host-parse") host_parse = true;
        else if (s.rfind("--", 0) == 0) test = s.substr(2);
    }
    if (geno.empty() || pheno.empty() || test.empty()) { usage(); return 1; }
    // --comp-level picks the reference's HOST layout (src/test/gwas_basic.cpp:265; genetics/genetic_data.cpp:60-79). The
    // device table replaces the bit-plane layouts 3 (2-bit blocks), 4 (three one-hot streams) and 5 (two streams), which
    // give identical counts; the byte-per-genotype tables 0-2 are not on this path.
    if (comp_level < 3 || comp_level > 5) {
        std::cerr << "--comp-level " << comp_level << ": only the bit-plane levels 3, 4 and 5 are replaced by the device table" << std::endl;
        return 1;
    }
    if (comp_level == 3 && (test == "test-boost-epi" || test == "contin-cc-perform")) {
        // the reference's level-3 table has no margins overloads: compressed_genotype_table3.h:250, .cpp:850-856 assert(false)
        std::cerr << "--comp-level 3 cannot run --" << test << ": the reference's 2-bit block table aborts in its margins overloads; use --comp-level 4 or 5" << std::endl;
        return 1;
    }
    if (devices < 1 || device + devices > gwasdev_device_count()) {
        std::cerr << "--devices " << devices << " from device " << device << ": " << gwasdev_device_count() << " CUDA devices visible" << std::endl;
        return 1;
    }

    std::set<int> cases, controls;
    int n_individs = 0;
    {
        std::ifstream f(pheno.c_str());
        if (!f.is_open()) { std::cerr << "cannot open " << pheno << std::endl; return 1; }
        std::string line;
        while (std::getline(f, line)) {
            if (line.empty()) continue;
            size_t pos = 0;
            for (int col = 0; col < 5 && pos != std::string::npos; ++col) pos = line.find_first_of("\t ", pos) == std::string::npos ? std::string::npos : line.find_first_of("\t ", pos) + 1;
            if (pos != std::string::npos && pos < line.size()) {
                if (line[pos] == '1') cases.insert(n_individs);
                else if (line[pos] == '0') controls.insert(n_individs);
            }
            ++n_individs;
        }
    }
    // Genotypes: the whole file is parsed, labelled and packed on the device (gwasdev_load_tped; .gz works, a .bed is
    // taken as SNP-major PLINK binary). --host-parse keeps the reference-shaped loop -- one addGenotypeRow per line
    // with the host packer -- as a cross-check.
    const bool is_bed = geno.size() > 4 && geno.compare(geno.size() - 4, 4, ".bed") == 0;
    std::cout << "Found " << n_individs << " individuals." << std::endl;
    if (!is_bed && !host_parse) {   // table sized from the file and filled in the same pass
        GeneticData gd_file(geno, device);
        if (gd_file.getGenotypedIndividualsCount() != n_individs) {
            std::cerr << geno << " has " << gd_file.getGenotypedIndividualsCount() << " genotype columns, " << pheno << " lists " << n_individs << " individuals" << std::endl;
            return 1;
        }
        std::cout << "Found " << gd_file.getGenotypedMarkersCount() << " markers" << std::endl;
        gd_file.getGenotypeTable()->useDevices(devices);
        return run_test(gd_file, cases, controls, test, outfile);
    }
    uint64_t n_rows64 = 0;
    uint32_t n_cols = 0;
    if (is_bed) {
        if (gwasdev_bed_dims(geno.c_str(), (uint32_t)n_individs, &n_rows64) != GWASDEV_OK) { std::cerr << gwasdev_last_error() << std::endl; return 1; }
    } else if (gwasdev_tped_dims(geno.c_str(), &n_rows64, &n_cols) != GWASDEV_OK) { std::cerr << gwasdev_last_error() << std::endl; return 1; }
    const int n_markers = (int)n_rows64;
    std::cout << "Found " << n_markers << " markers" << std::endl;
    GeneticData gd(n_markers, n_individs, device);
    if (is_bed) {
        // allele letters from the .bim next to the .bed (columns 5 and 6: A1, A2) decide the row header words, i.e. what
        // getCallAt() spells; rows whose alleles are not two different letters of ACGT keep the default A / C
        std::vector<unsigned char> alleles;
        std::ifstream bim((geno.substr(0, geno.size() - 4) + ".bim").c_str());
        if (bim.is_open()) {
            std::string chr, id, cm, pos, a1, a2;
            const std::string acgt = "ACGT";
            while (bim >> chr >> id >> cm >> pos >> a1 >> a2) {
                const size_t i1 = a1.size() == 1 ? acgt.find(a1[0]) : std::string::npos, i2 = a2.size() == 1 ? acgt.find(a2[0]) : std::string::npos;
                const bool ok = i1 != std::string::npos && i2 != std::string::npos && i1 != i2;
                alleles.push_back(ok ? (unsigned char)i1 : 0);
                alleles.push_back(ok ? (unsigned char)i2 : 1);
            }
            if ((int)alleles.size() != 2 * n_markers) alleles.clear();
        }
        gd.getGenotypeTable()->loadBed(geno, alleles);
    }
    else if (!host_parse) gd.getGenotypeTable()->loadTransposedPlink(geno);
    else {
        std::ifstream f(geno.c_str());
        std::string line, buf;
        int row = 0;
        while (std::getline(f, line)) {
            if (line.empty()) continue;
            size_t pos = 0;
            for (int col = 0; col < 4; ++col) pos = line.find_first_of("\t ", pos) + 1;
            buf.clear();
            bool second = false;
            for (size_t q = pos; q < line.size(); q += 2) {
                char c = line[q];
                c = c == '1' ? 'A' : c == '2' ? 'C' : c == '3' ? 'G' : c == '4' ? 'T' : c;
                buf.push_back(c);
                if (second) buf.push_back('\t');
                second = !second;
            }
            gd.addGenotypeRow(row++, buf.data(), buf.data() + buf.size(), '\t');
        }
    }
    gd.getGenotypeTable()->useDevices(devices);
    return run_test(gd, cases, controls, test, outfile);
}
