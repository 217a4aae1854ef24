// K8: the output lines, formatted on the device.
//
// Replaces the serial print loops of the reference -- matchAllImplementation.cpp:485-510 (one line per
// MatchPosAndError) and matchUniqueImplementation.cpp:252-321,1455-1486 (one line per read whose state is Straight or
// Reverse):
//
//     id \t bases \t [score] \t 1 \t a \t L \t +|- \t record name \t position in the record (1-based) \t \t k \n
//
// `bases` = toollib::remapString of the read, of its reverse complement for a '-' hit; the score goes through an ostream
// with default flags (printf's %g, csrc/fmt_g.h).  Two kernels per batch of items: k_fmt_len computes the length of every
// item's line (0: the item prints nothing), an exclusive scan turns the lengths into byte offsets, and k_fmt_write gives
// a warp to every line -- lane 0 lays out the short numeric pieces in shared memory, then the lanes copy the id and the
// record name byte by byte and expand the packed bases 32 at a time, so every store instruction of the warp writes
// consecutive bytes of the output.
#pragma once

#include "common.cuh"
#include "fmt_g.h"
#include "../../include/real_gpu.h"

namespace realgpu
{

struct FormatParams
{
        ReadSrc rs; const uint32_t * rlen;
        const char * ids; const uint64_t * id_off; uint64_t id_first;        // ids of the reads [id_first, id_first + ...)
        const uint32_t * file_first;                                         // [65] first record of every file in the tables below
        const char * names; const uint64_t * name_off; const uint64_t * rec_start;
        uint32_t scores;
        // items: reads [first, first+count) of the unique state, or rows [first, first+count) of the last matchAll
        const unsigned long long * info; const float * score;
        const real_gpu_hit * hits;
        uint64_t first, count;
        unsigned long long * nlines;  // items that print a line
        uint32_t * len;               // [count] bytes of every item's line
        const uint32_t * off;         // [count] exclusive scan of len
        char * out;
};

struct FormatItem
{
        uint64_t read, pos_in_record;
        uint32_t grec, k, inverted, L;          // grec: record index in the name tables
        float score;
        bool prints;
};

template<bool ALL>
__device__ __forceinline__ FormatItem format_item(FormatParams const & P, uint64_t i)
{
        FormatItem it;
        it.prints = false; it.read = 0; it.pos_in_record = 0; it.grec = 0; it.k = 0; it.inverted = 0; it.L = 0; it.score = 0;
        uint64_t pos; uint32_t file, frag;
        if ( ALL )
        {
                real_gpu_hit const h = P.hits[P.first + i];
                it.read = h.patid; pos = h.pos; file = h.file; frag = h.frag; it.k = h.k; it.inverted = h.inverted; it.score = h.score;
        }
        else
        {
                it.read = P.first + i;
                unsigned long long const d = P.info[it.read];
                uint32_t const st = umi_state(d);
                if ( st != ST_STRAIGHT && st != ST_REVERSE ) return it;
                pos = umi_pos(d); file = umi_file(d); frag = umi_frag(d); it.k = umi_err(d); it.inverted = st == ST_REVERSE;
                it.score = P.scores ? P.score[it.read] : 0.0f;
        }
        it.grec = P.file_first[file] + frag;
        it.pos_in_record = pos - P.rec_start[it.grec] + 1;
        it.L = P.rlen[it.read];
        it.prints = true;
        return it;
}

__device__ __forceinline__ uint32_t digits_u64(uint64_t v) { uint32_t n = 1; while ( v >= 10 ) { v /= 10; ++n; } return n; }

template<bool ALL>
__global__ void __launch_bounds__(256) k_fmt_len(FormatParams P)
{
        uint64_t const i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
        FormatItem it; it.prints = false;
        if ( i < P.count ) it = format_item<ALL>(P, i);
        uint32_t const bal = __ballot_sync(0xffffffffu, it.prints);
        if ( (threadIdx.x & 31) == 0 && bal ) atomicAdd(P.nlines, (unsigned long long)__popc(bal));
        if ( i >= P.count ) return;
        uint32_t n = 0;
        if ( it.prints )
        {
                uint64_t const r = it.read - P.id_first;
                char tmp[16];
                n = (uint32_t)(P.id_off[r+1] - P.id_off[r]) + 1 + it.L + 1 + (P.scores ? (uint32_t)fmtg::format_g6(it.score, tmp) : 0u) + 5 + digits_u64(it.L) + 3
                    + (uint32_t)(P.name_off[it.grec+1] - P.name_off[it.grec]) + 1 + digits_u64(it.pos_in_record) + 2 + digits_u64(it.k) + 1;
        }
        P.len[i] = n;
}

static const int FMT_WARPS = 8;

template<bool ALL>
__global__ void __launch_bounds__(FMT_WARPS * 32) k_fmt_write(FormatParams P)
{
        __shared__ char mid[FMT_WARPS][48], tail[FMT_WARPS][48];
        __shared__ uint32_t nmid[FMT_WARPS], ntail[FMT_WARPS];
        int const lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        for ( uint64_t i = (uint64_t)blockIdx.x * FMT_WARPS + wid; i < P.count; i += (uint64_t)gridDim.x * FMT_WARPS )
        {
                if ( ! P.len[i] ) continue;              // warp uniform
                FormatItem const it = format_item<ALL>(P, i);
                if ( lane == 0 )
                {
                        // \t score \t1\ta\t L \t+\t
                        char * m = mid[wid]; uint32_t n = 0;
                        m[n++] = '\t';
                        if ( P.scores ) n += (uint32_t)fmtg::format_g6(it.score, m + n);
                        m[n++] = '\t'; m[n++] = '1'; m[n++] = '\t'; m[n++] = 'a'; m[n++] = '\t';
                        n += (uint32_t)fmtg::format_u64(it.L, m + n);
                        m[n++] = '\t'; m[n++] = it.inverted ? '-' : '+'; m[n++] = '\t';
                        nmid[wid] = n;
                        // \t position \t\t k \n
                        char * t = tail[wid]; n = 0;
                        t[n++] = '\t';
                        n += (uint32_t)fmtg::format_u64(it.pos_in_record, t + n);
                        t[n++] = '\t'; t[n++] = '\t';
                        n += (uint32_t)fmtg::format_u64(it.k, t + n);
                        t[n++] = '\n';
                        ntail[wid] = n;
                }
                __syncwarp();
                char * o = P.out + P.off[i];
                uint64_t const r = it.read - P.id_first;
                const char * id = P.ids + P.id_off[r];
                uint32_t const idlen = (uint32_t)(P.id_off[r+1] - P.id_off[r]);
                for ( uint32_t b = lane; b < idlen; b += 32 ) o[b] = id[b];
                if ( lane == 0 ) o[idlen] = '\t';
                o += idlen + 1;
                // the bases of the strand that matched, 32 at a time (the loads are the same for every lane: one broadcast)
                uint32_t const sid = (uint32_t)it.read * 2 + it.inverted;
                for ( uint32_t w = 0; w * 32 < it.L; ++w )
                {
                        uint32_t const len = (it.L - 32*w < 32) ? (it.L - 32*w) : 32;
                        uint64_t const v = strand_bases(P.rs, sid, it.L, 32*w, len);
                        if ( (uint32_t)lane < len )
                                o[32*w + lane] = "ACGT"[(uint32_t)(v >> (62 - 2*lane)) & 3];
                }
                o += it.L;
                uint32_t const nm = nmid[wid], nt = ntail[wid];
                for ( uint32_t b = lane; b < nm; b += 32 ) o[b] = mid[wid][b];
                o += nm;
                const char * name = P.names + P.name_off[it.grec];
                uint32_t const namelen = (uint32_t)(P.name_off[it.grec+1] - P.name_off[it.grec]);
                for ( uint32_t b = lane; b < namelen; b += 32 ) o[b] = name[b];
                o += namelen;
                for ( uint32_t b = lane; b < nt; b += 32 ) o[b] = tail[wid][b];
                __syncwarp();
        }
}

// test hook: the score formatter on arbitrary floats; out[i] = 16 bytes, the characters followed by a zero byte
__global__ void __launch_bounds__(256) k_fmt_selftest(const float * __restrict__ v, uint64_t n, char * __restrict__ out)
{
        uint64_t const i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if ( i >= n ) return;
        char tmp[16];
        int const k = fmtg::format_g6(v[i], tmp);
        for ( int j = 0; j < 16; ++j ) out[i * 16 + j] = j < k ? tmp[j] : 0;
}

} // namespace realgpu
