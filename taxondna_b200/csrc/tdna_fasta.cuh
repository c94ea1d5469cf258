// FASTA bytes -> the n x L symbol matrix in the normalised alphabet, on the device.
//
// Restates, for the common case, FastaFile.appendFromFile + makeSequence (Common/DNA/formats/FastaFile.java:144-289: lines are
// blank / comment ('#' first) / name ('>' in column 0) / sequence; a record is a name line and the sequence lines up to the next
// name line, all whitespace removed, U -> T) and the alphabet part of Sequence.changeSequence (Common/DNA/Sequence.java:751-840:
// upper case, the leading and trailing runs of '-' -- '?' skipped over -- become '_', every symbol must pass isValid, a '_' in the
// input is an error).  The "[AG]" / "(AG)" group syntax (:758-796) is rare and sequential: such records are flagged and left to
// the host layer's normaliser.  HBM-bound byte work: every kernel reads the text or the line table once.
#pragma once
#include <cstdint>
// included after tdna_kernels.cuh (c_symlut: the pack kernel's symbol table)

namespace tdna {
namespace fasta {

enum : int { LINE_BLANK = 0, LINE_COMMENT = 1, LINE_NAME = 2, LINE_SEQ = 3 };
constexpr int BLOCK_BYTES = 4096;

__device__ __forceinline__ bool is_space(uint8_t ch) {  // java.util.regex \s: [ \t\n\x0B\f\r]
    return ch == ' ' || (ch >= 9 && ch <= 13);
}
// BufferedReader.readLine: a line ends at \n, \r or \r\n
__device__ __forceinline__ bool is_line_start(const uint8_t *__restrict__ text, int64_t i) {
    if (i == 0) return true;
    const uint8_t prev = text[i - 1];
    return prev == '\n' || (prev == '\r' && text[i] != '\n');
}

// 1. line starts per block of 4096 bytes
__global__ void __launch_bounds__(256) count_lines_kernel(const uint8_t *__restrict__ text, int64_t nbytes, int *__restrict__ block_lines) {
    __shared__ int s_cnt;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * BLOCK_BYTES;
    int c = 0;
    for (int k = threadIdx.x; k < BLOCK_BYTES; k += 256) {
        const int64_t i = base + k;
        if (i < nbytes && is_line_start(text, i)) c++;
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) c += __shfl_xor_sync(0xffffffffu, c, m);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s_cnt, c);
    __syncthreads();
    if (threadIdx.x == 0) block_lines[blockIdx.x] = s_cnt;
}

// 2. the line table, in order: one warp per block walks its bytes 32 at a time (ballot + popc keeps the order)
__global__ void __launch_bounds__(32) list_lines_kernel(const uint8_t *__restrict__ text, int64_t nbytes, const long long *__restrict__ block_first,
                                                        long long *__restrict__ line_start) {
    const int64_t base = (int64_t)blockIdx.x * BLOCK_BYTES;
    long long w = block_first[blockIdx.x];
    const int lane = threadIdx.x;
    for (int k = 0; k < BLOCK_BYTES; k += 32) {
        const int64_t i = base + k + lane;
        const bool st = i < nbytes && is_line_start(text, i);
        const unsigned m = __ballot_sync(0xffffffffu, st);
        if (st) line_start[w + __popc(m & ((1u << lane) - 1u))] = i;
        w += __popc(m);
    }
}

// 3. per line: its kind and how many sequence symbols (non-whitespace bytes) it carries
__global__ void __launch_bounds__(256) classify_lines_kernel(const uint8_t *__restrict__ text, int64_t nbytes, const long long *__restrict__ line_start,
                                                             int64_t nlines, int *__restrict__ kind, int *__restrict__ symbols, int *__restrict__ is_name) {
    const int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= nlines) return;
    const int64_t b = line_start[l], e = l + 1 < nlines ? line_start[l + 1] : nbytes;
    int k = LINE_BLANK, cnt = 0;
    if (text[b] == '>') k = LINE_NAME;  // "^>": column 0 only
    else {
        for (int64_t i = b; i < e; i++) {
            const uint8_t ch = text[i];
            if (is_space(ch)) continue;
            if (k == LINE_BLANK) k = ch == '#' ? LINE_COMMENT : LINE_SEQ;
            if (k == LINE_COMMENT) break;
            cnt++;
        }
    }
    kind[l] = k;
    symbols[l] = k == LINE_SEQ ? cnt : 0;
    is_name[l] = k == LINE_NAME ? 1 : 0;
}

// 4. per name line: the record's header span and where its symbols start in the running symbol count
__global__ void __launch_bounds__(256) records_kernel(const uint8_t *__restrict__ text, int64_t nbytes, const long long *__restrict__ line_start, int64_t nlines,
                                                      const int *__restrict__ kind, const long long *__restrict__ rec_of_line,
                                                      const long long *__restrict__ sym_before_line, long long *__restrict__ rec_line,
                                                      long long *__restrict__ header_offset, int *__restrict__ header_length) {
    const int64_t l = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= nlines || kind[l] != LINE_NAME) return;
    const long long r = rec_of_line[l];  // exclusive count of name lines before this one
    rec_line[r] = l;
    const int64_t b = line_start[l] + 1;
    int64_t e = l + 1 < nlines ? line_start[l + 1] : nbytes;
    while (e > b && (text[e - 1] == '\n' || text[e - 1] == '\r')) e--;
    header_offset[r] = b;
    header_length[r] = (int)(e - b);
}
__global__ void __launch_bounds__(256) record_lengths_kernel(const long long *__restrict__ rec_line, int64_t nrec, int64_t nlines,
                                                             const long long *__restrict__ sym_before_line, int *__restrict__ seq_length,
                                                             int *__restrict__ max_length) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrec) return;
    const long long l0 = rec_line[r], l1 = r + 1 < nrec ? rec_line[r + 1] : nlines;
    const long long len = sym_before_line[l1] - sym_before_line[l0];
    seq_length[r] = (int)len;
    atomicMax(max_length, (int)len);
}

// 5. the symbols of every sequence line into its record's row: upper case, U -> T; one warp per line
enum : int { FLAG_URACIL = 1, FLAG_INVALID = 2, FLAG_NEEDS_HOST = 4 };
__global__ void __launch_bounds__(256) scatter_symbols_kernel(const uint8_t *__restrict__ text, int64_t nbytes, const long long *__restrict__ line_start,
                                                              int64_t nlines, const int *__restrict__ kind, const long long *__restrict__ rec_of_line,
                                                              const long long *__restrict__ sym_before_line, const long long *__restrict__ rec_line,
                                                              uint8_t *__restrict__ rows, int64_t row_stride,
                                                              int *__restrict__ flags) {
    const int64_t l = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (l >= nlines || kind[l] != LINE_SEQ) return;
    const long long r = rec_of_line[l] - 1;  // name lines strictly before this line, minus one
    if (r < 0) return;                       // sequence lines before the first name are dropped (FastaFile.java:186-203)
    long long w = sym_before_line[l] - sym_before_line[rec_line[r]];
    const int64_t b = line_start[l], e = l + 1 < nlines ? line_start[l + 1] : nbytes;
    uint8_t *dst = rows + r * row_stride;
    int fl = 0;
    for (int64_t i0 = b; i0 < e; i0 += 32) {
        const int64_t i = i0 + lane;
        uint8_t ch = i < e ? text[i] : (uint8_t)' ';
        const bool sym = !is_space(ch);
        const unsigned m = __ballot_sync(0xffffffffu, sym);
        if (sym) {
            if (ch >= 'a' && ch <= 'z') ch = (uint8_t)(ch - 32);  // String.toUpperCase on ASCII
            if (ch == 'U') { ch = 'T'; fl |= FLAG_URACIL; }
            if (ch == '(' || ch == '[' || ch == ')' || ch == ']') fl |= FLAG_NEEDS_HOST;
            else if (c_symlut[ch] & 0x8000u) fl |= FLAG_INVALID;  // fails isValid (:927-965); '_' passes it -- the strokes below reject it
            dst[w + __popc(m & ((1u << lane) - 1u))] = ch;
        }
        w += __popc(m);
    }
    if (fl) atomicOr(flags + r, fl);
}

// 6. per record: external gaps (the leading and the trailing run of '-', '?' skipped over) -> '_', the row's tail -> '?'.  One warp per row.
__global__ void __launch_bounds__(256) external_gaps_kernel(uint8_t *__restrict__ rows, int64_t row_stride, int width, const int *__restrict__ seq_length,
                                                            int64_t nrec, int *__restrict__ flags) {
    const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= nrec) return;
    uint8_t *row = rows + r * row_stride;
    const int len = seq_length[r];
    // first and last symbol that is neither '-' nor '?'
    int first = len, last = -1;
    for (int i = lane; i < len; i += 32) {
        const uint8_t ch = row[i];
        if (ch != '-' && ch != '?') { first = min(first, i); last = max(last, i); }
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
        first = min(first, __shfl_xor_sync(0xffffffffu, first, m));
        last = max(last, __shfl_xor_sync(0xffffffffu, last, m));
    }
    // a '_' in the input is only an error where a stroke meets it, i.e. as the first or the last symbol that is neither '-' nor '?'
    // (:812-819, :828-838); anywhere else it passes isValid and stays
    if (lane == 0 && last >= 0 && (row[first] == '_' || row[last] == '_')) atomicOr(flags + r, FLAG_INVALID);
    // nothing but gaps and '?': both strokes together turn every '-' into '_' (the backward stroke stops where the forward one ended)
    for (int i = lane; i < width; i += 32) {
        if (i >= len) row[i] = '?';
        else if (row[i] == '-' && (i < first || i > last)) row[i] = '_';
    }
}

// ---- consensus of a bin (Cluster.makeConsensusOfBin, SpeciesIdentifier/Cluster.java:123-138) --------------------------------
// The reference folds Sequence.getConsensus (Common/DNA/Sequence.java:1374-1398; per column consensus(), :1008-1050) over the
// members, re-normalising after every step.  Column by column the fold has a closed form: '?' is absorbing (a member that has
// ended counts as '?'), letters combine as the union of their IUPAC sets, gaps of either kind give way to letters -- so the result
// is '?' if any member has '?' there, else the code of the union if any member has a letter, else a gap; which gap ('-' or
// '_') is decided by the final normalisation alone (external_gaps_kernel).  One CTA per bin, threads over columns.
__global__ void __launch_bounds__(256) consensus_kernel(const uint8_t *__restrict__ sym, const int32_t *__restrict__ len, int64_t sym_stride, int L,
                                                        const int32_t *__restrict__ members, const long long *__restrict__ offsets,
                                                        uint8_t *__restrict__ out, int64_t out_stride, int *__restrict__ out_len) {
    // IUPAC code of a set of bases, bits A C G T as in the pack kernel's table (getcode, Sequence.java:1149-1196, with its A C T G bits)
    const char code[16] = {'-', 'A', 'C', 'M', 'G', 'R', 'S', 'V', 'T', 'W', 'Y', 'H', 'K', 'D', 'B', 'N'};
    const long long b = offsets[blockIdx.x], e = offsets[blockIdx.x + 1];
    int maxlen = 0;
    for (long long k = b; k < e; k++) maxlen = max(maxlen, min(len[members[k]], L));
    if (threadIdx.x == 0) out_len[blockIdx.x] = maxlen;
    uint8_t *dst = out + (int64_t)blockIdx.x * out_stride;
    for (int col = threadIdx.x; col < maxlen; col += 256) {
        bool missing = false;
        uint32_t bases = 0;
        for (long long k = b; k < e; k++) {
            const int row = members[k];
            if (col >= len[row]) { missing = true; break; }
            const uint8_t ch = sym[(int64_t)row * sym_stride + col];
            if (ch == '?') { missing = true; break; }
            bases |= c_symlut[ch] & 15u;
        }
        dst[col] = missing ? (uint8_t)'?' : (uint8_t)code[bases];
    }
}

}  // namespace fasta
}  // namespace tdna
