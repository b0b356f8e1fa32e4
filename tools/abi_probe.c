#include <stddef.h>
#include "dsv.h"
#include "dsv_encoder.h"
#include "dsv_decoder.h"
#define S(t) printf("sizeof %s %zu\n", #t, sizeof(t))
#define O(t,f) printf("offsetof %s.%s %zu\n", #t, #f, offsetof(t,f))
int main(void){
 S(DSV_META);S(DSV_PLANE);S(DSV_COEFS);S(DSV_FRAME);S(DSV_MV);S(DSV_PARAMS);S(DSV_BUF);S(DSV_ENCODER);S(DSV_DECODER);S(DSV_HME);
 O(DSV_PLANE,data);O(DSV_PLANE,len);O(DSV_PLANE,format);O(DSV_PLANE,stride);O(DSV_PLANE,w);O(DSV_PLANE,h);O(DSV_PLANE,hs);O(DSV_PLANE,vs);
 O(DSV_FRAME,alloc);O(DSV_FRAME,planes);O(DSV_FRAME,refcount);O(DSV_FRAME,format);O(DSV_FRAME,width);O(DSV_FRAME,height);O(DSV_FRAME,border);
 O(DSV_MV,u);O(DSV_MV,mode);O(DSV_MV,submask);O(DSV_MV,lo_var);O(DSV_MV,lo_tex);O(DSV_MV,high_detail);
 O(DSV_ENCODER,quality);O(DSV_ENCODER,gop);O(DSV_ENCODER,do_scd);O(DSV_ENCODER,rc_mode);O(DSV_ENCODER,rc_high_motion_nudge);O(DSV_ENCODER,bitrate);
 O(DSV_ENCODER,max_q_step);O(DSV_ENCODER,min_quality);O(DSV_ENCODER,max_quality);O(DSV_ENCODER,min_I_frame_quality);O(DSV_ENCODER,intra_pct_thresh);
 O(DSV_ENCODER,scene_change_delta);O(DSV_ENCODER,stable_refresh);O(DSV_ENCODER,pyramid_levels);O(DSV_ENCODER,rc_quant);O(DSV_ENCODER,next_fnum);
 O(DSV_ENCODER,ref);O(DSV_ENCODER,vidmeta);O(DSV_ENCODER,prev_link);O(DSV_ENCODER,force_metadata);O(DSV_ENCODER,stability);O(DSV_ENCODER,refresh_ctr);
 O(DSV_ENCODER,stable_blocks);O(DSV_ENCODER,prev_gop);O(DSV_ENCODER,prev_avg_luma);
 O(DSV_DECODER,vidmeta);O(DSV_DECODER,ref);O(DSV_DECODER,draw_info);O(DSV_DECODER,got_metadata);
 O(DSV_PARAMS,vidmeta);O(DSV_PARAMS,is_ref);O(DSV_PARAMS,has_ref);O(DSV_PARAMS,blk_w);O(DSV_PARAMS,blk_h);O(DSV_PARAMS,nblocks_h);O(DSV_PARAMS,nblocks_v);
 O(DSV_HME,params);O(DSV_HME,src);O(DSV_HME,ref);O(DSV_HME,mvf);O(DSV_HME,levels);
 O(DSV_BUF,data);O(DSV_BUF,len);O(DSV_META,width);O(DSV_META,aspect_den);
 return 0;}
