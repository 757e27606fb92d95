// C++ host-mirror test: the reference's example tests (align/sw_example_test.go,
// nw_example_test.go, fitted_example_test.go, align_test.go:28-51) written against
// include/biogo_align.hpp.  Linked against whichever build of the C ABI the test passes
// (the CPU-emulated kernels in the no-GPU tier, the CUDA library on the GPU box).
#include <cstdio>
#include <cstring>

#include "biogo_align.hpp"

using namespace biogo;

static align::Linear match(int gap, int m, int x)
{
    align::Linear M(5, std::vector<int>(5, x));
    for (int i = 0; i < 5; i++)
        for (int j = 0; j < 5; j++) M[i][j] = (i == 0 || j == 0) ? gap : (i == j ? m : x);
    M[0][0] = 0;
    return M;
}

static int fails = 0;
#define CHECK(cond)                                              \
    do {                                                         \
        if (!(cond)) {                                           \
            printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #cond); \
            fails++;                                             \
        }                                                        \
    } while (0)

int main()
{
    const alphabet::Alphabet *dna = &alphabet::DNAgapped();
    std::vector<feat::Pair> aln;
    const align::Linear nw4 = {{0, -5, -5, -5, -5}, {-5, 10, -3, -1, -4}, {-5, -3, 9, -5, 0}, {-5, -1, -5, 7, -3}, {-5, -4, 0, -3, 8}};

    {   // ExampleSW_Align_a
        align::SW smith{match(-1, 2, -1)};
        auto a = seq::Seq(dna, "ACACACTA"), b = seq::Seq(dna, "AGCACACA");
        CHECK(smith.Align(a, b, aln) == align::nil);
        CHECK(feat::Sprint(aln) == "[[0,1)/[0,1)=2 -/[1,2)=-1 [1,6)/[2,7)=10 [6,7)/-=-1 [7,8)/[7,8)=2]");
        auto fa = align::Format(a, b, aln, '-');
        CHECK(fa[0] == "A-CACACTA" && fa[1] == "AGCACAC-A");
    }
    {   // ExampleSWAffine_Align, also through QLetters
        align::SWAffine smith{match(-1, 1, -1), -5};
        auto a = seq::Seq(dna, "ATAGGAAG"), b = seq::Seq(dna, "ATTGGCAATG");
        CHECK(smith.Align(a, b, aln) == align::nil);
        CHECK(feat::Sprint(aln) == "[[0,7)/[0,7)=3]");
        auto qa = seq::QSeq(dna, "ATAGGAAG"), qb = seq::QSeq(dna, "ATTGGCAATG");
        CHECK(smith.Align(qa, qb, aln) == align::nil);
        CHECK(feat::Sprint(aln) == "[[0,7)/[0,7)=3]");
    }
    {   // ExampleNW_Align / ExampleNWAffine_Align / TestIssue58
        align::NW needle{nw4};
        auto a = seq::Seq(dna, "AGACTAGTTA"), b = seq::Seq(dna, "GACAGACG");
        CHECK(needle.Align(a, b, aln) == align::nil);
        CHECK(feat::Sprint(aln) == "[[0,1)/-=-5 [1,4)/[0,3)=26 [4,5)/-=-5 [5,10)/[3,8)=12]");
        auto fa = align::Format(a, b, aln, '-');
        CHECK(fa[0] == "AGACTAGTTA" && fa[1] == "-GAC-AGACG");
        align::NWAffine na{match(-1, 1, -1), -5};
        CHECK(na.Align(seq::Seq(dna, "ATAGGAAG"), seq::Seq(dna, "ATTGGCAATG"), aln) == align::nil);
        CHECK(feat::Sprint(aln) == "[[0,7)/[0,7)=3 -/[7,9)=-7 [7,8)/[9,10)=1]");
        align::NWAffine n58{match(-1, 1, -1), -1};
        CHECK(n58.Align(seq::Seq(dna, "GCTCACTAAAAACACAATCTACAACAGACGTTGCACTAACACTGTAATTGCCTTTAGTCC"),
                        seq::Seq(dna, "ACTGCGTA"), aln) == align::nil);
        CHECK(feat::Sprint(aln) ==
              "[[0,4)/-=-5 [4,7)/[0,3)=3 [7,32)/-=-26 [32,34)/[3,5)=2 [34,43)/-=-10 [43,46)/[5,8)=3 [46,60)/-=-15]");
    }
    {   // ExampleFitted_Align / ExampleFittedAffine_Align
        align::Fitted fit{nw4};
        auto a = seq::Seq(dna, "GTTGACAGACTAGATTCACG"), b = seq::Seq(dna, "GACAGACGA");
        CHECK(fit.Align(a, b, aln) == align::nil);
        CHECK(feat::Sprint(aln) == "[[3,10)/[0,7)=62 [10,12)/-=-10 [12,14)/[7,9)=17]");
        align::FittedAffine fa{match(-1, 1, -1), -5};
        CHECK(fa.Align(seq::Seq(dna, "ATTGGCAATGA"), seq::Seq(dna, "ATAGGAA"), aln) == align::nil);
        CHECK(feat::Sprint(aln) == "[[0,7)/[0,7)=3]");
    }
    {   // error behaviour (align/sw.go:23-48, align/sw_letters.go:69-79,97-102)
        align::SW smith{match(-1, 2, -1)};
        CHECK(smith.Align(seq::Seq(nullptr, "ACGT"), seq::Seq(dna, "ACGT"), aln) == align::ErrNoAlphabet);
        CHECK(smith.Align(seq::Seq(dna, "ACGT"), seq::Seq(&alphabet::DNAredundant(), "ACGT"), aln) == align::ErrMismatchedAlphabets);
        CHECK(smith.Align(seq::Seq(&alphabet::DNA(), "ACGT"), seq::Seq(&alphabet::DNA(), "ACGT"), aln) == align::ErrNotGappedAlphabet);
        CHECK(smith.Align(seq::Seq(dna, "ACGT"), seq::QSeq(dna, "ACGT"), aln) == align::ErrMismatchedTypes);
        align::SW small{{{0, -1}, {-1, 1}}};
        CHECK(small.Align(seq::Seq(dna, "ACGT"), seq::Seq(dna, "ACGT"), aln) == align::ErrMatrixWrongSize(2, 5));
        align::Error e = smith.Align(seq::Seq(dna, "ACGN"), seq::Seq(dna, "AXGT"), aln);
        CHECK(e.What() == "align: illegal letter 'X' at position 1 in qSeq");
        bool panicked = false;
        try {
            align::NWAffine{match(-1, 1, -1), -2}.Align(seq::Seq(dna, ""), seq::Seq(dna, "ACG"), aln);
        } catch (const align::GoPanic &) {
            panicked = true;
        }
        CHECK(panicked);
    }
    printf(fails ? "mirror_test: %d FAILED\n" : "mirror_test: all ok\n", fails);
    return fails ? 1 : 0;
}
