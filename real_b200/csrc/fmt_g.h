// The score column of an output line: a float printed through an ostream with default flags
// (matchAllImplementation.cpp:497, matchUniqueImplementation.cpp:279), i.e. like printf("%g"): six significant
// digits, correctly rounded (half to even on the exact binary value, as glibc does), trailing zeros dropped,
// exponent form when the decimal exponent is below -4 or above 5.
//
// Exact integer arithmetic only -- a float is m * 2^e with m < 2^24, so the six digits are
// round(m * 2^e * 10^(5-X)) for the decimal exponent X of the value; 64-bit integers cover every value between
// 1e-7 and 2^63, a small multi-word integer the rest of the float range.  Compiles for the host too (the CPU
// test-suite checks it against snprintf on millions of bit patterns).
#pragma once

#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define FMTG_HD __host__ __device__
#else
#define FMTG_HD
#endif

namespace fmtg
{

static const int BIG_LIMBS = 7;                 // 224 bits: m * 10^50 < 2^24 * 2^167, m * 2^104 < 2^128

struct Big
{
        uint32_t w[BIG_LIMBS];                  // little endian
};

FMTG_HD inline void big_set(Big & b, uint64_t v)
{
        for ( int i = 0; i < BIG_LIMBS; ++i ) b.w[i] = 0;
        b.w[0] = (uint32_t)v; b.w[1] = (uint32_t)(v >> 32);
}
FMTG_HD inline void big_mul_small(Big & b, uint32_t f)
{
        uint64_t carry = 0;
        for ( int i = 0; i < BIG_LIMBS; ++i )
        {
                uint64_t const t = (uint64_t)b.w[i] * f + carry;
                b.w[i] = (uint32_t)t; carry = t >> 32;
        }
}
// b /= d, returns the remainder
FMTG_HD inline uint32_t big_div_small(Big & b, uint32_t d)
{
        uint64_t rem = 0;
        for ( int i = BIG_LIMBS - 1; i >= 0; --i )
        {
                uint64_t const t = (rem << 32) | b.w[i];
                b.w[i] = (uint32_t)(t / d); rem = t % d;
        }
        return (uint32_t)rem;
}
FMTG_HD inline void big_shl(Big & b, uint32_t s)
{
        uint32_t const ws = s >> 5, bs = s & 31;
        for ( int i = BIG_LIMBS - 1; i >= 0; --i )
        {
                uint32_t v = 0;
                if ( i >= (int)ws )
                {
                        v = b.w[i - ws] << bs;
                        if ( bs && i >= (int)ws + 1 ) v |= b.w[i - ws - 1] >> (32 - bs);
                }
                b.w[i] = v;
        }
}
FMTG_HD inline bool big_bit(Big const & b, uint32_t i) { return i < 32u * BIG_LIMBS && ((b.w[i >> 5] >> (i & 31)) & 1u); }
// any bit below position s set?
FMTG_HD inline bool big_any_below(Big const & b, uint32_t s)
{
        for ( uint32_t i = 0; i < (uint32_t)BIG_LIMBS; ++i )
        {
                if ( 32 * i >= s ) break;
                uint32_t const m = (s - 32 * i >= 32) ? 0xFFFFFFFFu : ((1u << (s - 32 * i)) - 1);
                if ( b.w[i] & m ) return true;
        }
        return false;
}
// b >>= s
FMTG_HD inline void big_shr(Big & b, uint32_t s)
{
        uint32_t const ws = s >> 5, bs = s & 31;
        for ( int i = 0; i < BIG_LIMBS; ++i )
        {
                uint32_t v = 0;
                if ( i + (int)ws < BIG_LIMBS )
                {
                        v = b.w[i + ws] >> bs;
                        if ( bs && i + (int)ws + 1 < BIG_LIMBS ) v |= b.w[i + ws + 1] << (32 - bs);
                }
                b.w[i] = v;
        }
}
FMTG_HD inline uint64_t big_low64(Big const & b) { return ((uint64_t)b.w[1] << 32) | b.w[0]; }
FMTG_HD inline bool big_fits64(Big const & b) { for ( int i = 2; i < BIG_LIMBS; ++i ) if ( b.w[i] ) return false; return true; }

// round-half-even(m * 2^e * 10^p) for any p; returns false when the result does not fit 64 bits
FMTG_HD inline bool scaled_round(uint32_t m, int e, int p, uint64_t & out)
{
        Big b;
        big_set(b, m);
        if ( p > 0 )
        {
                int k = p;
                while ( k >= 9 ) { big_mul_small(b, 1000000000u); k -= 9; }
                uint32_t f = 1;
                for ( int i = 0; i < k; ++i ) f *= 10u;
                if ( k ) big_mul_small(b, f);
        }
        if ( e > 0 ) big_shl(b, (uint32_t)e);
        // now the value is b * 2^min(e,0) / 10^max(-p,0)
        bool sticky = false;            // something non-zero was dropped below the current last digit
        uint32_t half_num = 0, half_den = 1;   // the fraction dropped last, as half_num / half_den (exact when ! sticky)
        bool have_frac = false;
        if ( e < 0 )
        {
                uint32_t const s = (uint32_t)(-e);
                bool const hb = big_bit(b, s - 1);
                bool const below = s >= 2 && big_any_below(b, s - 1);
                big_shr(b, s);
                // fraction = (hb ? 1/2 : 0) + something below 1/2 when `below`
                have_frac = true;
                if ( p < 0 )
                {
                        // a further division follows: everything dropped so far is below one unit of b -- only its presence matters
                        sticky = hb || below;
                        have_frac = false;
                }
                else
                {
                        half_num = hb ? 1u : 0u; half_den = 2;
                        sticky = below;
                }
        }
        if ( p < 0 )
        {
                int k = -p;
                uint32_t r = 0;
                while ( k > 1 ) { r = big_div_small(b, 10u); if ( r ) sticky = true; --k; }
                r = big_div_small(b, 10u);
                have_frac = true; half_num = r; half_den = 10;
        }
        if ( ! big_fits64(b) ) return false;
        uint64_t v = big_low64(b);
        if ( have_frac )
        {
                // compare half_num/half_den (+ sticky) with 1/2
                uint32_t const twice = 2 * half_num;
                bool up = false;
                if ( twice > half_den ) up = true;
                else if ( twice == half_den ) up = sticky || (v & 1);
                if ( up ) ++v;
        }
        out = v;
        return true;
}

// the characters of printf("%g", (double)x); returns their number (at most 13: "-1.23457e-38"); no terminator
FMTG_HD inline int format_g6(float x, char * o)
{
        uint32_t bits;
        memcpy(&bits, &x, 4);
        int n = 0;
        if ( bits >> 31 ) o[n++] = '-';
        uint32_t const ex = (bits >> 23) & 0xFF, fr = bits & 0x7FFFFF;
        if ( ex == 0xFF )
        {
                if ( fr ) { o[n++] = 'n'; o[n++] = 'a'; o[n++] = 'n'; }
                else { o[n++] = 'i'; o[n++] = 'n'; o[n++] = 'f'; }
                return n;
        }
        if ( ex == 0 && fr == 0 ) { o[n++] = '0'; return n; }
        uint32_t const m = ex ? (fr | 0x800000u) : fr;
        int const e = ex ? (int)ex - 150 : -149;              // |x| = m * 2^e
        // decimal exponent: estimate from the position of the leading bit, then correct
        int msb = 31; while ( ! ((m >> msb) & 1u) ) --msb;
        int X = (int)(((long long)(msb + e) * 78913LL) >> 18);        // floor((msb+e) * log10(2)) for this range, at most one too small
        uint64_t D = 0;
        for ( int tries = 0; tries < 6; ++tries )
        {
                // D = 10^6: the value was rounded up to 10^(X+1) (or lies just above it); one exponent up it reads 100000
                if ( ! scaled_round(m, e, 5 - X, D) || D >= 1000000ULL ) { ++X; continue; }
                if ( D < 100000ULL ) { --X; continue; }
                break;
        }
        char dg[6];
        { uint64_t t = D; for ( int i = 5; i >= 0; --i ) { dg[i] = (char)('0' + (int)(t % 10)); t /= 10; } }
        int last = 5; while ( last > 0 && dg[last] == '0' ) --last;          // last significant digit
        if ( X < -4 || X >= 6 )
        {
                o[n++] = dg[0];
                if ( last > 0 ) { o[n++] = '.'; for ( int i = 1; i <= last; ++i ) o[n++] = dg[i]; }
                o[n++] = 'e';
                int ax = X;
                if ( ax < 0 ) { o[n++] = '-'; ax = -ax; } else o[n++] = '+';
                o[n++] = (char)('0' + ax / 10); o[n++] = (char)('0' + ax % 10);     // |X| <= 45
                return n;
        }
        if ( X >= 0 )
        {
                for ( int i = 0; i <= X; ++i ) o[n++] = dg[i];
                if ( last > X ) { o[n++] = '.'; for ( int i = X + 1; i <= last; ++i ) o[n++] = dg[i]; }
                return n;
        }
        o[n++] = '0'; o[n++] = '.';
        for ( int i = 0; i < -X - 1; ++i ) o[n++] = '0';
        for ( int i = 0; i <= last; ++i ) o[n++] = dg[i];
        return n;
}

// decimal digits of v; returns their number (at most 20)
FMTG_HD inline int format_u64(uint64_t v, char * o)
{
        char tmp[20]; int k = 0;
        do { tmp[k++] = (char)('0' + (int)(v % 10)); v /= 10; } while ( v );
        for ( int i = 0; i < k; ++i ) o[i] = tmp[k - 1 - i];
        return k;
}

} // namespace fmtg
