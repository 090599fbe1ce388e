// fastx.hpp -- FASTA / FASTQ records as `count` consumes them (count.go:283-330 iterates fastx records: the sequence with
// line breaks and blanks removed, case kept): concatenated bases + record offsets, the layout ukm_count_seq takes.
//
// Host code above the C ABI (the CLI's reader); unit-tested on the CPU by tests/host/fastx_test.cpp.
//   FASTA: '>' header line, then sequence lines until the next line that starts with '>'.
//   FASTQ: '@' header line, sequence lines until a line that starts with '+', then quality lines until as many quality
//          characters as bases have been read (so a quality line may start with '@' or '>', and records may be wrapped).
// Blanks (space, tab, CR) inside sequence lines are dropped.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <string>
#include <vector>

namespace fastx {

// Appends the records of `raw` to bases / rec_off (rec_off gets one END offset per record; the caller seeds it with 0).
// Returns "" or an error message.
inline std::string parse(const uint8_t* raw, size_t n, std::vector<uint8_t>& bases, std::vector<uint64_t>& rec_off) {
    size_t i = 0;
    auto line_end = [&](size_t p) {
        while (p < n && raw[p] != '\n') ++p;
        return p;
    };
    auto blank = [](uint8_t c) { return c == ' ' || c == '\t' || c == '\r'; };
    auto append_line = [&](size_t b, size_t e) {
        for (size_t p = b; p < e; ++p)
            if (!blank(raw[p])) bases.push_back(raw[p]);
    };
    while (i < n) {
        if (raw[i] == '\n' || blank(raw[i])) { ++i; continue; }
        if (raw[i] == '>') {
            i = line_end(i) + 1;
            while (i < n && raw[i] != '>') {
                const size_t e = line_end(i);
                append_line(i, e);
                i = e + 1;
            }
            rec_off.push_back(bases.size());
        } else if (raw[i] == '@') {
            i = line_end(i) + 1;
            const size_t start = bases.size();
            while (i < n && raw[i] != '+') {
                const size_t e = line_end(i);
                append_line(i, e);
                i = e + 1;
            }
            if (i >= n) return "truncated FASTQ record (no '+' line)";
            rec_off.push_back(bases.size());
            i = line_end(i) + 1;  // the '+' line
            const size_t want = bases.size() - start;
            size_t have = 0;
            while (i < n && have < want) {
                const size_t e = line_end(i);
                for (size_t p = i; p < e; ++p)
                    if (!blank(raw[p])) ++have;
                i = e + 1;
            }
            if (have != want) return "FASTQ record: quality length differs from sequence length";
        } else {
            return "invalid FASTA/Q record start";
        }
    }
    return "";
}

}  // namespace fastx
