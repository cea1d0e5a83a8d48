// placeholder, replaced below
#include "rb_internal.h"
rb_status rb_lbvh_build(rb_raycaster *) { return rb_fail(RB_ERR_UNSUPPORTED, "lbvh: not built yet"); }
rb_status rb_lbvh_refit(rb_raycaster *, const float *, size_t) { return rb_fail(RB_ERR_UNSUPPORTED, "lbvh: not built yet"); }
rb_status rb_lbvh_download(const rb_raycaster *, rb_node *, rb_tri *, float *) { return rb_fail(RB_ERR_UNSUPPORTED, "lbvh: not built yet"); }
