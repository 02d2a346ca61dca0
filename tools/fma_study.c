#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#define PI 3.141592653589793
static double rhsF(double u,double M3){ return fma(M3*u,u,-u);} 
static double rhsS(double u,double M3){ return -u + M3*u*u; }
static int orbit(int fused,double M,double R_S,double r_obs,double alpha,double *phi_o,double*u_o,double*w_o,int*steps_o){
  double f0=1.0-R_S/r_obs; double b=r_obs*sin(alpha)/sqrt(f0); if(b==0.0) return 0;
  double u=1.0/r_obs; double w0=1.0/(b*b)-u*u+2.0*M*u*u*u; if(w0<0) return 0; double w=sqrt(w0);
  double phi=0,uc=1.0/(R_S*1.01),ue=1.0/(2.0*r_obs); int status=2; int steps=0; double h=0.05,hh=0.5*h,h6=h/6.0,M3=3.0*M;
  while(phi<50.0){ double rem=50.0-phi; double hs=h; if(rem<hs){hs=rem;} if(hs<=0)break; 
    double up=u,wp=w; 
    if(fused && hs==h){
      double k1u=w,k1w=rhsF(u,M3); double ua=fma(hh,k1u,u),wa=fma(hh,k1w,w);
      double k2u=wa,k2w=rhsF(ua,M3); double ub=fma(hh,k2u,u),wb=fma(hh,k2w,w);
      double k3u=wb,k3w=rhsF(ub,M3); double ucc=fma(h,k3u,u),wc=fma(h,k3w,w);
      double k4u=wc,k4w=rhsF(ucc,M3);
      double su=fma(2.0,k3u,fma(2.0,k2u,k1u))+k4u, sw=fma(2.0,k3w,fma(2.0,k2w,k1w))+k4w;
      u=fma(h6,su,up); w=fma(h6,sw,wp);
    } else {
      double H=hs,HH=0.5*hs,H6=hs/6.0;
      double k1u=w,k1w=rhsS(u,M3); double ua=up+HH*k1u,wa=wp+HH*k1w;
      double k2u=wa,k2w=rhsS(ua,M3); double ub=up+HH*k2u,wb=wp+HH*k2w;
      double k3u=wb,k3w=rhsS(ub,M3); double ucc=up+H*k3u,wc=wp+H*k3w;
      double k4u=wc,k4w=rhsS(ucc,M3);
      u=up+H6*(k1u+2.0*k2u+2.0*k3u+k4u); w=wp+H6*(k1w+2.0*k2w+2.0*k3w+k4w);
    }
    steps++;
    if(up<uc&&u>=uc){double d=u-up;double fr=d==0?1:(uc-up)/d; if(fr<0)fr=0;if(fr>1)fr=1; phi+=fr*hs; w=wp+fr*(w-wp);u=uc;status=-1;break;}
    if(up>ue&&u<=ue){double d=u-up;double fr=d==0?1:(ue-up)/d; if(fr<0)fr=0;if(fr>1)fr=1; phi+=fr*hs; w=wp+fr*(w-wp);u=ue;status=1;break;}
    phi+=hs; }
  *phi_o=phi;*u_o=u;*w_o=w;*steps_o=steps; return status; }
static int ray(int fused,double M,double r_obs,double alpha,double*fa,long*nh,int*steps){
  double phi,u,w; int st=orbit(fused,M,2*M,r_obs,alpha,&phi,&u,&w,steps); if(st==0){*fa=NAN;*nh=0;return 0;}
  double r=1.0/u; *nh=(long)floor(fabs(phi)/PI); if(st==-1||r<=2.2*M){*fa=NAN;return -1;}
  double dr=-w/(u*u); double hy=dr*sin(phi)+r*cos(phi),hx=dr*cos(phi)-r*sin(phi); double c=-cos(atan2(hy,hx)); if(c>1)c=1;if(c<-1)c=-1; *fa=acos(c); return 1;}
int main(int argc,char**argv){
  double robs[]={15,25,50,100,300,1000}; 
  for(int ir=0;ir<6;ir++){ double r_obs=robs[ir]; double M=1; double ac=asin(3*sqrt(3.0)*sqrt(1-2/r_obs)/r_obs);
    double maxrel[40]={0}; long cnt[40]={0}; long flips=0, nhdiff=0; int flipminsteps=100000; 
    // scan alpha: log-spaced offsets around ac plus uniform
    long N=4000000; 
    for(long i=0;i<N;i++){ double t=(double)i/N; double alpha; 
      if(i%2==0){ double e=pow(10,-14+13.5*t); alpha=ac*(1+((i/2)%2?e:-e)); } else alpha=ac*(0.2+3.0*t);
      double fa1,fa2; long n1,n2; int s1,s2; int st1=ray(0,M,r_obs,alpha,&fa1,&n1,&s1); int st2=ray(1,M,r_obs,alpha,&fa2,&n2,&s2);
      int bin=s2/25; if(bin>39)bin=39; cnt[bin]++;
      if(st1!=st2){flips++; if(s2<flipminsteps)flipminsteps=s2;}
      else { if(n1!=n2){nhdiff++; if(s2<flipminsteps) flipminsteps=s2;} if(st1==1){ double d=fabs(fa1-fa2)/fmax(fa1,1e-3); if(d>maxrel[bin])maxrel[bin]=d; } }
    }
    printf("r_obs=%g flips=%ld nhdiff=%ld min steps(fused) of any flip=%d\n",r_obs,flips,nhdiff,flipminsteps);
    for(int b=0;b<40;b++) if(cnt[b]) printf("  steps %4d-%4d n=%8ld maxrel=%.3e\n",b*25,b*25+24,cnt[b],maxrel[b]);
  }
}
