#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <omp.h>
static inline uint64_t sm64(uint64_t* s){ uint64_t z=(*s+=0x9E3779B97F4A7C15ULL); z=(z^(z>>30))*0xBF58476D1CE4E5B9ULL; z=(z^(z>>27))*0x94D049BB133111EBULL; return z^(z>>31);}
int main(){
  const double y = 1.0/6.0; // RN(1/6)
  long long bad=0, total=0;
  #pragma omp parallel reduction(+:bad,total)
  {
    uint64_t s = 12345 + 7919*omp_get_thread_num();
    for (long long it=0; it<400000000LL; it++){
      uint64_t bits = sm64(&s);
      // exponent restricted to [-900, 900]
      uint64_t mant = bits & 0xFFFFFFFFFFFFFULL;
      int e = (int)((bits>>52)&0x7FF);
      e = 64 + (e % 1920);
      uint64_t b = ((bits>>63)<<63) | ((uint64_t)e<<52) | mant;
      double a; memcpy(&a,&b,8);
      double q = a*y;
      double r = fma(-6.0, q, a);
      double q2 = fma(r, y, q);
      double ref = a/6.0;
      if (q2 != ref) bad++;
      total++;
    }
    // structured hard cases: mantissas of the form (6k + small)/..., all exponents
    for (uint64_t m=0; m< (1ULL<<22); m++){
      for (int sh=0; sh<31; sh+=5){
        uint64_t mant = (m<<sh) & 0xFFFFFFFFFFFFFULL;
        uint64_t b = ((uint64_t)1023<<52) | mant;
        double a; memcpy(&a,&b,8);
        if (omp_get_thread_num()!=0) break;
        double q = a*y; double r = fma(-6.0,q,a); double q2=fma(r,y,q);
        if (q2 != a/6.0) bad++;
        total++;
      }
    }
  }
  printf("total %lld mismatches %lld\n", total, bad);
  return 0;
}
