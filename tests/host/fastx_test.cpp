// fastx_test.cpp -- CPU unit test of the CLI's FASTA / FASTQ reader (unikmer_b200/host/fastx.hpp).
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../unikmer_b200/host/fastx.hpp"

static int g_fail = 0;

static void expect(const char* name, const std::string& text, const std::vector<std::string>& recs, bool want_error = false) {
    std::vector<uint8_t> bases;
    std::vector<uint64_t> off{0};
    const std::string err = fastx::parse(reinterpret_cast<const uint8_t*>(text.data()), text.size(), bases, off);
    bool ok = want_error ? !err.empty() : err.empty();
    if (ok && !want_error) {
        ok = off.size() == recs.size() + 1;
        for (size_t r = 0; ok && r < recs.size(); ++r)
            ok = std::string(bases.begin() + off[r], bases.begin() + off[r + 1]) == recs[r];
    }
    if (!ok) {
        fprintf(stderr, "FAIL %s (err='%s', %zu records)\n", name, err.c_str(), off.size() - 1);
        ++g_fail;
    }
}

int main() {
    expect("fasta wrapped", ">a desc\nACGT\nacgtn\n>b\nGG\n", {"ACGTacgtn", "GG"});
    expect("fasta crlf, blanks inside, no final newline", ">a\r\nAC GT\r\nAC\tGT \r\n\r\n>b\r\nTT", {"ACGTACGT", "TT"});
    expect("fasta empty record", ">a\n>b\nAC\n", {"", "AC"});
    expect("fastq 4-line", "@r1\nACGT\n+\nIIII\n@r2\nGG\n+r2\n##\n", {"ACGT", "GG"});
    expect("fastq wrapped, quality lines starting with @ and >", "@r1\nACGT\nACGT\n+\n@III\n>III\n@r2\nTTT\n+\n@@@\n", {"ACGTACGT", "TTT"});
    expect("fastq crlf", "@r1\r\nAC\r\n+\r\nII\r\n", {"AC"});
    expect("leading blank lines", "\n\r\n>a\nAC\n\n>b\nTT\n\n", {"AC", "TT"});
    expect("empty input", "", {});
    expect("garbage start", "ACGT\n", {}, true);
    expect("fastq truncated", "@r1\nACGT\n", {}, true);
    expect("fastq short quality", "@r1\nACGT\n+\nII\n", {}, true);
    if (g_fail) return 1;
    printf("fastx ok\n");
    return 0;
}
