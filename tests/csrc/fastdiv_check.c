#include <stdio.h>
#include "rt_fastdiv.h"
int main(void){
  uint32_t ds[]={1,2,3,5,7,8,16,60,64,100,127,128,240,480,1000,1024,4096,65535};
  uint32_t nm[]={100,65535,1u<<20,(1u<<28)-1,(1u<<29)+12345,0xffffffffu};
  for(unsigned i=0;i<sizeof ds/4;i++)for(unsigned j=0;j<sizeof nm/4;j++){
    RT_FastDiv f=rt_fastdiv_make(ds[i],nm[j]);
    uint64_t step = nm[j] > (1u<<24) ? 997 : 1;
    for(uint64_t n=0;n<=nm[j];n+=step){ if(rt_fastdiv((uint32_t)n,f,ds[i])!=(uint32_t)n/ds[i]){printf("FAIL d=%u n=%llu\n",ds[i],(unsigned long long)n);return 1;} }
    // boundaries
    for(uint64_t q=0;q<=nm[j]/ds[i];q+= (nm[j]/ds[i]/100000+1)){ uint64_t n=q*ds[i]; if(n<=nm[j] && rt_fastdiv((uint32_t)n,f,ds[i])!=q){printf("FAILb d=%u n=%llu\n",ds[i],(unsigned long long)n);return 1;} if(n&&rt_fastdiv((uint32_t)(n-1),f,ds[i])!=q-1){printf("FAILc\n");return 1;} }
    if(j==3) printf("d=%u nmax=%u mul=%u shift=%u\n",ds[i],nm[j],f.mul,f.shift);
  }
  printf("ok\n");return 0;}
