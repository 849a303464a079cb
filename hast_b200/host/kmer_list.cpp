// kmer_list.cpp -- read a parent-unique k-mer list (one k-mer per line, the
// output of 00.build_unshare_kmers_by_jellyfish/build_unshared_kmers.sh:290-291)
// with the line semantics of load_kmers, classify.cpp:30-46:
//   * k is the length of the FIRST line of the hap0 file (:35-36), taken even
//     when that line is the unterminated last one (:35-39);
//   * afterwards the loop stops at the first getline that reaches end-of-file,
//     so a last line without '\n' is dropped (:41);
//   * a line of any other length trips assert(str.size()==overlap), kmer.h:154.
// Only framing is decided here.  Letters are packed, canonicalised and checked
// on the device (hast_table_add_text), straight from this text.
#include <sys/stat.h>

#include <cstdio>
#include <cstring>

#include "host.h"

namespace hasthost {

std::string load_kmer_list(const std::string& path, int index, int k_in, KmerList& out) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return "cannot open k-mer list " + path;
    struct stat st;
    if (fstat(fileno(f), &st) != 0) { fclose(f); return "cannot stat " + path; }
    std::string& text = out.text;
    text.resize((size_t)st.st_size);
    size_t got = text.empty() ? 0 : fread(&text[0], 1, text.size(), f);
    fclose(f);
    if (got != text.size()) return "short read on " + path;

    int k = k_in;
    if (index == 0) {
        const char* nl = (const char*)memchr(text.data(), '\n', text.size());
        const size_t len = nl ? (size_t)(nl - text.data()) : text.size();
        if (len < 1 || len > 32)
            return "k = " + std::to_string(len) + " (length of the first line of " + path +
                   ") is outside 1..32; the reference is only correct for k <= 32 (kmer.h:225-238)";
        k = (int)len;
        if (!nl) text.push_back('\n');                    // the lone unterminated first line still counts
    }
    out.k = k;
    const size_t stride = (size_t)k + 1;
    const size_t n_lines = text.size() / stride;
    const size_t tail = text.size() - n_lines * stride;
    // an unterminated tail is dropped; a terminated one is a line of the wrong length
    if (tail && memchr(text.data() + n_lines * stride, '\n', tail))
        return "k-mer line of length != " + std::to_string(k) + " near the end of " + path +
               " (the reference asserts, kmer.h:154)";
    text.resize(n_lines * stride);
    out.n_lines = n_lines;
    return "";
}

}  // namespace hasthost
