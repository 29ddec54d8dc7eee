/*
 * ORACLE (test infrastructure, NOT product code) -- type-generic body.
 *
 * Included twice by geometry_ops.c with
 *     T      = float  / double
 *     FN(x)  = x##_f32 / x##_f64
 * Every function restates, loop for loop, the CPU kernel of one lmbspecialops
 * op of the reference (lmb-freiburg/demon, submodule lmbspecialops @ fb5004d).
 * The file:line of the code being restated is cited on each function.
 *
 * Third-party arithmetic that is NOT under /root/reference: the reference
 * builds on Eigen (the copy bundled with the TensorFlow 1.4.0 pip package,
 * Eigen 3.3.x).  Where the reference calls into Eigen (AngleAxis /
 * Quaternion::toRotationMatrix, fixed-size dot products, JacobiSVD::solve)
 * we restate Eigen's published algorithm in plain C and say so at the site.
 *
 * Compile with -ffp-contract=off: the reference is x86-64 code without FMA
 * contraction, every a*b+c below is two IEEE roundings.
 */

/* x86 cvttss2si / cvttsd2si semantics of `(int)x` (warp2d.cc:196 does
 * `p2.template cast<int>()`): truncation toward zero, and the "integer
 * indefinite" value INT_MIN for NaN and for anything outside int range.
 * Doing it by hand keeps the oracle free of C undefined behaviour. */
static int FN(cvtt)(T x)
{
  if (!(x > (T)-2147483649.0 && x < (T)2147483648.0)) return INT_MIN;
  return (int)x;
}

/* ------------------------------------------------------------------------
 * warp2d  --  lmbspecialops/src/warp2d.cc:171-256 (Warp2dOp::warp2d_cpu)
 * border_mode: 1 = CLAMP, 2 = VALUE (enum at warp2d.cc:259)
 * in  [w][z][y][x], displacements [w][2][y][x], out like in.
 * The 4-term dot product is Eigen's fixed-size Vec4 dot; with SSE packets it
 * reduces as (p0+p2)+(p1+p3) (predux of a 4-lane packet / of two 2-lane
 * packets), which is what is restated here.
 * ---------------------------------------------------------------------- */
void FN(oracle_warp2d)(T* out, const T* in, const T* displacements,
                       int x_size, int y_size, int z_size, int w_size,
                       int normalized, int border_mode, T border_value)
{
  const long xy_size = (long)x_size * y_size;
  const long xyz_size = xy_size * z_size;
#define IN_(w,z,y,x) in[(w)*xyz_size+(z)*xy_size+(long)(y)*x_size+(x)]
#define OUT_(w,z,y,x) out[(w)*xyz_size+(z)*xy_size+(long)(y)*x_size+(x)]
#define VEC_(w,z,y,x) displacements[(w)*2*xy_size+(z)*xy_size+(long)(y)*x_size+(x)]
  for (int w = 0; w < w_size; ++w)
    for (int y = 0; y < y_size; ++y)
      for (int x = 0; x < x_size; ++x) {
        T p1x = (T)x, p1y = (T)y;
        T vx = VEC_(w,0,y,x), vy = VEC_(w,1,y,x);
        if (normalized) { vx *= x_size; vy *= y_size; }
        T p2x = p1x + vx, p2y = p1y + vy;
        int p2ix = FN(cvtt)(p2x), p2iy = FN(cvtt)(p2y);
        T a = p2x - (T)p2ix;
        T b = p2y - (T)p2iy;
        T w0 = ((T)1-a)*((T)1-b), w1 = a*((T)1-b), w2 = ((T)1-a)*b, w3 = a*b;
        /* unsigned wrap mirrors the two's complement wrap of INT_MIN+1 etc. */
        int x0, y0, x1, y1, x2, y2, x3, y3;
        int px1 = (int)((unsigned)p2ix + 1u), py1 = (int)((unsigned)p2iy + 1u);
        if (border_mode == 1) {
#define CL_(v,n) ((v) < 0 ? 0 : ((v) > (n)-1 ? (n)-1 : (v)))
          x0 = CL_(p2ix, x_size); y0 = CL_(p2iy, y_size);
          x1 = CL_(px1,  x_size); y1 = CL_(p2iy, y_size);
          x2 = CL_(p2ix, x_size); y2 = CL_(py1,  y_size);
          x3 = CL_(px1,  x_size); y3 = CL_(py1,  y_size);
#undef CL_
          for (int z = 0; z < z_size; ++z) {
            T v0 = IN_(w,z,y0,x0), v1 = IN_(w,z,y1,x1), v2 = IN_(w,z,y2,x2), v3 = IN_(w,z,y3,x3);
            OUT_(w,z,y,x) = (v0*w0 + v2*w2) + (v1*w1 + v3*w3);
          }
        } else {
          x0 = p2ix; y0 = p2iy; x1 = px1; y1 = p2iy; x2 = p2ix; y2 = py1; x3 = px1; y3 = py1;
          for (int z = 0; z < z_size; ++z) {
            if (x0 >= 0 && x3 > 0 && x3 < x_size && y0 >= 0 && y3 > 0 && y3 < y_size) {
              T v0 = IN_(w,z,y0,x0), v1 = IN_(w,z,y1,x1), v2 = IN_(w,z,y2,x2), v3 = IN_(w,z,y3,x3);
              OUT_(w,z,y,x) = (v0*w0 + v2*w2) + (v1*w1 + v3*w3);
            } else {
              OUT_(w,z,y,x) = border_value;
            }
          }
        }
      }
#undef IN_
#undef OUT_
#undef VEC_
}

/* ------------------------------------------------------------------------
 * rotation formats  --  lmbspecialops/src/rotation_format.h:38-82
 * format: 0 = MATRIX (row major 3x3), 1 = QUATERNION (w,x,y,z), 2 = ANGLEAXIS3
 * R is written row major.  AngleAxis::toRotationMatrix and
 * Quaternion::toRotationMatrix are Eigen 3.3 (third party, absent from
 * /root/reference); their published algorithms are restated.
 * ---------------------------------------------------------------------- */
static int FN(rotation_step)(int format) { return format == 0 ? 9 : (format == 1 ? 4 : 3); }

static void FN(to_rotation_matrix)(T* R, const T* data, int format)
{
  if (format == 0) {
    for (int i = 0; i < 9; ++i) R[i] = data[i];
  } else if (format == 1) {
    T w = data[0], x = data[1], y = data[2], z = data[3];
    /* q.normalize(): coeffs /= norm */
    T n = FN(sqrt_)(((x*x + y*y) + z*z) + w*w);
    w /= n; x /= n; y /= n; z /= n;
    T tx = (T)2*x, ty = (T)2*y, tz = (T)2*z;
    T twx = tx*w, twy = ty*w, twz = tz*w;
    T txx = tx*x, txy = ty*x, txz = tz*x;
    T tyy = ty*y, tyz = tz*y, tzz = tz*z;
    R[0] = (T)1-(tyy+tzz); R[1] = txy-twz;        R[2] = txz+twy;
    R[3] = txy+twz;        R[4] = (T)1-(txx+tzz); R[5] = tyz-twx;
    R[6] = txz-twy;        R[7] = tyz+twx;        R[8] = (T)1-(txx+tyy);
  } else {
    T ax = data[0], ay = data[1], az = data[2];
    T angle = FN(sqrt_)((ax*ax + ay*ay) + az*az);
    if (angle > (T)1.0e-6) {
      ax /= angle; ay /= angle; az /= angle;
      T s = FN(sin_)(angle), c = FN(cos_)(angle);
      T sx = s*ax, sy = s*ay, sz = s*az;
      T c1x = ((T)1-c)*ax, c1y = ((T)1-c)*ay, c1z = ((T)1-c)*az;
      T tmp;
      tmp = c1x*ay; R[1] = tmp - sz; R[3] = tmp + sz;
      tmp = c1x*az; R[2] = tmp + sy; R[6] = tmp - sy;
      tmp = c1y*az; R[5] = tmp - sx; R[7] = tmp + sx;
      R[0] = c1x*ax + c; R[4] = c1y*ay + c; R[8] = c1z*az + c;
    } else {
      R[0]=1; R[1]=0; R[2]=0; R[3]=0; R[4]=1; R[5]=0; R[6]=0; R[7]=0; R[8]=1;
    }
  }
}

void FN(oracle_rotation_matrix)(T* R, const T* data, int format, int n)
{
  for (int i = 0; i < n; ++i) FN(to_rotation_matrix)(R + 9*i, data + i*FN(rotation_step)(format), format);
}

/* ------------------------------------------------------------------------
 * depth_to_flow -- lmbspecialops/src/depthtoflow.cc:250-313 (depthtoflow_cpu)
 *                  and compute_flow, depthtoflow.cc:158-185
 * depth [z][y][x], intrinsics [z][4], rotation [z][step], translation [z][3]
 * out [z][2][y][x]
 * ---------------------------------------------------------------------- */
void FN(oracle_depth_to_flow)(T* out, const T* depth, const T* intrinsics,
                              const T* rotation, const T* translation,
                              int x_size, int y_size, int z_size,
                              int rotation_format, int inverse_depth, int normalize_flow)
{
  const long xy_size = (long)x_size * y_size;
  const T inv_x_size = (T)(1.0 / x_size);   /* depthtoflow.cc:261 computes in double, then casts */
  const T inv_y_size = (T)(1.0 / y_size);
  const int step = FN(rotation_step)(rotation_format);
  for (int z = 0; z < z_size; ++z) {
    T fx = intrinsics[4*z+0]*x_size, fy = intrinsics[4*z+1]*y_size;
    T cx = intrinsics[4*z+2]*x_size, cy = intrinsics[4*z+3]*y_size;
    T inv_fx = (T)1/fx, inv_fy = (T)1/fy;
    const T* t = translation + 3*z;
    T R[9];
    FN(to_rotation_matrix)(R, rotation + (long)z*step, rotation_format);
    const T* depthmap = depth + z*xy_size;
    T* flow = out + 2*z*xy_size;
    for (int y = 0; y < y_size; ++y)
      for (int x = 0; x < x_size; ++x) {
        T fvx, fvy;
        T d = depthmap[(long)y*x_size + x];
        if (inverse_depth) d = (T)1/d;
        if (d > 0 && FN(isfinite_)(d)) {
          T p1x = x + (T)0.5, p1y = y + (T)0.5;
          /* compute_flow */
          T t2x = (p1x - cx)*inv_fx, t2y = (p1y - cy)*inv_fy;
          T X0 = d*t2x, X1 = d*t2y, X2 = d*(T)1;
          T p2x = ((R[0]*X0 + R[1]*X1) + R[2]*X2) + t[0];
          T p2y = ((R[3]*X0 + R[4]*X1) + R[5]*X2) + t[1];
          T p2z = ((R[6]*X0 + R[7]*X1) + R[8]*X2) + t[2];
          p2x = fx*(p2x/p2z) + cx;
          p2y = fy*(p2y/p2z) + cy;
          fvx = p2x - p1x; fvy = p2y - p1y;
          if (normalize_flow) { fvx *= inv_x_size; fvy *= inv_y_size; }
        } else {
          fvx = FN(nan_)(); fvy = FN(nan_)();
        }
        flow[(long)y*x_size + x] = fvx;
        flow[xy_size + (long)y*x_size + x] = fvy;
      }
  }
}

/* ------------------------------------------------------------------------
 * Least squares  X = argmin |A X - b|, A 4x3  --  stands in for
 * Eigen::JacobiSVD<Matrix<T,Dynamic,3>>(A, FullU|FullV).solve(b)
 * (flowtodepth.cc:275-278).  Eigen is absent; restated with the one-sided
 * (Hestenes) Jacobi SVD in precision T and Eigen's rank rule
 * (SVDBase::rank(): singular values <= max(1,diagSize)*eps*sigma_max are
 * treated as zero, diagSize = 3).
 * ---------------------------------------------------------------------- */
static void FN(svd_solve_4x3)(T* X, const T A_in[4][3], const T b[4])
{
  T U[4][3], V[3][3] = {{1,0,0},{0,1,0},{0,0,1}};
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 3; ++j) U[i][j] = A_in[i][j];
  const T eps = FN(eps_)();
  for (int sweep = 0; sweep < 60; ++sweep) {
    int rotated = 0;
    for (int p = 0; p < 2; ++p)
      for (int q = p+1; q < 3; ++q) {
        T alpha = 0, beta = 0, gamma = 0;
        for (int i = 0; i < 4; ++i) { alpha += U[i][p]*U[i][p]; beta += U[i][q]*U[i][q]; gamma += U[i][p]*U[i][q]; }
        if (gamma == 0 || FN(abs_)(gamma) <= eps*FN(sqrt_)(alpha*beta)) continue;
        rotated = 1;
        T zeta = (beta - alpha)/((T)2*gamma);
        T tt = (zeta >= 0 ? (T)1 : (T)-1)/(FN(abs_)(zeta) + FN(sqrt_)((T)1 + zeta*zeta));
        T c = (T)1/FN(sqrt_)((T)1 + tt*tt), s = c*tt;
        for (int i = 0; i < 4; ++i) { T up = U[i][p], uq = U[i][q]; U[i][p] = c*up - s*uq; U[i][q] = s*up + c*uq; }
        for (int i = 0; i < 3; ++i) { T vp = V[i][p], vq = V[i][q]; V[i][p] = c*vp - s*vq; V[i][q] = s*vp + c*vq; }
      }
    if (!rotated) break;
  }
  T sig2[3], smax2 = 0;
  for (int j = 0; j < 3; ++j) { sig2[j] = 0; for (int i = 0; i < 4; ++i) sig2[j] += U[i][j]*U[i][j]; if (sig2[j] > smax2) smax2 = sig2[j]; }
  const T thr = (T)3*eps*FN(sqrt_)(smax2);
  X[0] = X[1] = X[2] = 0;
  for (int j = 0; j < 3; ++j) {
    T sigma = FN(sqrt_)(sig2[j]);
    if (!(sigma > thr)) { if (sigma != sigma) { X[0] = X[1] = X[2] = sigma; return; } continue; }
    T ub = 0; for (int i = 0; i < 4; ++i) ub += U[i][j]*b[i];
    T coef = ub / sig2[j];
    for (int i = 0; i < 3; ++i) X[i] += coef*V[i][j];
  }
}

/* ------------------------------------------------------------------------
 * flow_to_depth -- lmbspecialops/src/flowtodepth.cc:383-481 (flowtodepth_cpu)
 *                  triangulateLinear, flowtodepth.cc:251-281
 * The fundamental matrix / epipolar projection (flowtodepth.cc:207-248,
 * 452-458) is dead code: xvec[1] = x2, not x2_on_line (flowtodepth.cc:461-462);
 * it is not restated.  flowtodepth2.cc differs only inside that dead code, so
 * this function is the oracle for flow_to_depth AND flow_to_depth2.
 * flow [z][2][y][x] -> out [z][y][x]
 * ---------------------------------------------------------------------- */
void FN(oracle_flow_to_depth)(T* out, const T* flow, const T* intrinsics,
                              const T* rotation, const T* translation,
                              int x_size, int y_size, int z_size,
                              int rotation_format, int inverse_depth, int normalized_flow)
{
  const long xy_size = (long)x_size * y_size;
  const T inv_x_size = (T)(1.0 / x_size);
  const T inv_y_size = (T)(1.0 / y_size);
  const int step = FN(rotation_step)(rotation_format);
  for (int z = 0; z < z_size; ++z) {
    T K[9] = {intrinsics[4*z+0], 0, intrinsics[4*z+2],  0, intrinsics[4*z+1], intrinsics[4*z+3],  0, 0, 1};
    const T* t = translation + 3*z;
    T R[9];
    FN(to_rotation_matrix)(R, rotation + (long)z*step, rotation_format);
    T P1[3][4], Rt[3][4], P2[3][4];
    for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) { P1[i][j] = K[3*i+j]; Rt[i][j] = R[3*i+j]; } P1[i][3] = 0; Rt[i][3] = t[i]; }
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 4; ++j)
      P2[i][j] = (K[3*i+0]*Rt[0][j] + K[3*i+1]*Rt[1][j]) + K[3*i+2]*Rt[2][j];
    T* depthmap = out + z*xy_size;
    const T* flowmap = flow + 2*z*xy_size;
    for (int y = 0; y < y_size; ++y)
      for (int x = 0; x < x_size; ++x) {
        T x1x = (x + (T)0.5)*inv_x_size, x1y = (y + (T)0.5)*inv_y_size;
        T fx_ = flowmap[(long)y*x_size + x], fy_ = flowmap[xy_size + (long)y*x_size + x];
        if (!normalized_flow) { fx_ *= inv_x_size; fy_ *= inv_y_size; }
        T x2x = x1x + fx_, x2y = x1y + fy_;
        T A[4][3], b[4];
        const T px[2] = {x1x, x2x}, py[2] = {x1y, x2y};
        for (int i = 0; i < 2; ++i) {
          const T (*P)[4] = (i == 0) ? P1 : P2;
          for (int j = 0; j < 3; ++j) {
            A[2*i+0][j] = py[i]*P[2][j] - (T)1*P[1][j];
            A[2*i+1][j] = (T)1*P[0][j] - px[i]*P[2][j];
          }
          b[2*i+0] = (T)1*P[1][3] - py[i]*P[2][3];
          b[2*i+1] = px[i]*P[2][3] - (T)1*P[0][3];
        }
        T X[3];
        FN(svd_solve_4x3)(X, A, b);
        if (FN(isfinite_)(X[0]) && FN(isfinite_)(X[1]) && FN(isfinite_)(X[2]) && X[2] > 0)
          depthmap[(long)y*x_size + x] = inverse_depth ? (T)1/X[2] : X[2];
        else
          depthmap[(long)y*x_size + x] = 0;
      }
  }
}

/* ------------------------------------------------------------------------
 * leaky_relu -- lmbspecialops/src/leakyrelu.cc:62-82   out = std::max(leak*x, x)
 * std::max(a,b) = (a < b) ? b : a
 * ---------------------------------------------------------------------- */
void FN(oracle_leaky_relu)(T* out, const T* in, long size, T leak)
{
  for (long i = 0; i < size; ++i) { T tmp = in[i]; T a = leak*tmp; out[i] = (a < tmp) ? tmp : a; }
}

/* ------------------------------------------------------------------------
 * median3x3_downsample -- lmbspecialops/src/median3x3downsample.cc:112-184
 * in [z][in_y][in_x] -> out [z][ceil(in_y/2)][ceil(in_x/2)]
 * Five passes "bubble the minimum of v[k..8] into v[k]" with strict '>'.
 * ---------------------------------------------------------------------- */
void FN(oracle_median3x3_downsample)(T* out, const T* in, long z_size, int in_y_size, int in_x_size)
{
  T* out_ptr = out;
  for (long z = 0; z < z_size; ++z)
    for (int y = 0; y < in_y_size; y += 2)
      for (int x = 0; x < in_x_size; x += 2) {
        T value[9];
        int idx = 0;
        for (int dy = -1; dy <= 1; ++dy)
          for (int dx = -1; dx <= 1; ++dx) {
            int x_ = x+dx; if (x_ < 0) x_ = 0; if (x_ > in_x_size-1) x_ = in_x_size-1;
            int y_ = y+dy; if (y_ < 0) y_ = 0; if (y_ > in_y_size-1) y_ = in_y_size-1;
            value[idx++] = in[z*in_y_size*in_x_size + (long)y_*in_x_size + x_];
          }
        for (int k = 0; k < 5; ++k)
          for (int j = k+1; j < 9; ++j)
            if (value[k] > value[j]) { T tmp = value[k]; value[k] = value[j]; value[j] = tmp; }
        *out_ptr++ = value[4];
      }
}

/* ------------------------------------------------------------------------
 * scale_invariant_gradient (forward) --
 *   lmbspecialops/src/scaleinvariantgradient.cc:148-195 (scaleinvariantgrad_cpu)
 * in [z][y][x] -> out [z][2][y][x]
 * ---------------------------------------------------------------------- */
void FN(oracle_scale_invariant_gradient)(T* out, const T* in, int x_size, int y_size, long z_size,
                                         const int* deltas, const T* weights, int num, T eps)
{
  const long xy_size = (long)x_size * y_size;
  for (long z = 0; z < z_size; ++z) {
    T* out_z = out + 2*z*xy_size;
    const T* in_z = in + z*xy_size;
    for (int y = 0; y < y_size; ++y)
      for (int x = 0; x < x_size; ++x) {
        const T value0 = in_z[(long)y*x_size + x];
        T grad_x = 0, grad_y = 0;
        for (int c = 0; c < num; ++c) {
          int delta = deltas[c];
          T weight = weights[c];
          T valuex = (x+delta >= 0 && x+delta < x_size) ? in_z[(long)y*x_size + x+delta] : value0;
          T valuey = (y+delta >= 0 && y+delta < y_size) ? in_z[(long)(y+delta)*x_size + x] : value0;
          grad_x += weight*(valuex-value0)/((FN(abs_)(value0)+FN(abs_)(valuex))+eps);
          grad_y += weight*(valuey-value0)/((FN(abs_)(value0)+FN(abs_)(valuey))+eps);
        }
        out_z[(long)y*x_size + x] = grad_x;
        out_z[xy_size + (long)y*x_size + x] = grad_y;
      }
  }
}

/* ------------------------------------------------------------------------
 * depth_to_normals --
 *   lmbspecialops/src/depthtonormals.cc:147-238 (depthtonormals_cpu), compute3dPoint :95-101
 * depth [z][y][x], intrinsics [z][4] -> out [z][3][y][x]; border pixels and pixels with a
 * non-positive / non-finite depth among their 4-neighbourhood are NaN.
 * Eigen pieces restated (Eigen 3.3, THIRD PARTY, absent here):
 *   Matrix3::inverse()   Eigen/src/LU/InverseImpl.h compute_inverse<.,.,3>: cofactors of column 0,
 *                        det = (c0*m00 + c1*m10) + c2*m20, every entry = cofactor * (1/det)
 *   cross(), normalize() (z = squaredNorm(); if (z > 0) v /= sqrt(z)), left-to-right sums
 * ---------------------------------------------------------------------- */
static void FN(d2n_point_)(T p[3], int x, int y, T depth, T i00, T i02, T i11, T i12)
{
  p[0] = ((x + (T)0.5)*i00 + i02)*depth;
  p[1] = ((y + (T)0.5)*i11 + i12)*depth;
  p[2] = depth;
}
static void FN(d2n_cross_)(T r[3], const T a[3], const T b[3])
{
  r[0] = a[1]*b[2] - a[2]*b[1];
  r[1] = a[2]*b[0] - a[0]*b[2];
  r[2] = a[0]*b[1] - a[1]*b[0];
}
static void FN(d2n_normalize_)(T v[3])
{
  const T z = (v[0]*v[0] + v[1]*v[1]) + v[2]*v[2];
  if (z > (T)0) {
    const T n = FN(sqrt_)(z);
    v[0] = v[0]/n; v[1] = v[1]/n; v[2] = v[2]/n;
  }
}

void FN(oracle_depth_to_normals)(T* out, const T* depth, const T* intrinsics, int x_size, int y_size, long z_size, int inverse_depth)
{
  const long xy_size = (long)x_size * y_size;
  for (long z = 0; z < z_size; ++z) {
    /* K = [[a,0,cx],[0,b,cy],[0,0,1]] */
    const T a = intrinsics[4*z+0]*x_size, b = intrinsics[4*z+1]*y_size, cx = intrinsics[4*z+2]*x_size, cy = intrinsics[4*z+3]*y_size;
    const T c0 = b*(T)1 - cy*(T)0;
    const T c1 = (T)0*cx - (T)1*(T)0;
    const T c2 = (T)0*cy - cx*b;
    const T det = (c0*a + c1*(T)0) + c2*(T)0;
    const T invdet = (T)1/det;
    const T i00 = c0*invdet, i02 = c2*invdet;
    const T i11 = ((T)1*a - (T)0*cx)*invdet;
    const T i12 = (cx*(T)0 - a*cy)*invdet;
    const T* dm = depth + z*xy_size;
    T* normal = out + 3*z*xy_size;
    for (int y = 0; y < y_size; ++y)
      for (int x = 0; x < x_size; ++x) {
        T n0 = FN(nan_)(), n1 = n0, n2 = n0;
        if (!(x == 0 || y == 0 || x == x_size-1 || y == y_size-1)) {
          T d = dm[(long)y*x_size + x], d_y0 = dm[(long)(y-1)*x_size + x], d_x0 = dm[(long)y*x_size + x-1];
          T d_y1 = dm[(long)(y+1)*x_size + x], d_x1 = dm[(long)y*x_size + x+1];
          if (inverse_depth) { d = 1/d; d_y0 = 1/d_y0; d_x0 = 1/d_x0; d_y1 = 1/d_y1; d_x1 = 1/d_x1; }
          if (!(d <= 0 || !FN(isfinite_)(d) || d_y0 <= 0 || !FN(isfinite_)(d_y0) || d_x0 <= 0 || !FN(isfinite_)(d_x0) ||
                d_y1 <= 0 || !FN(isfinite_)(d_y1) || d_x1 <= 0 || !FN(isfinite_)(d_x1))) {
            T p[3], p_y0[3], p_x0[3], p_y1[3], p_x1[3], a1[3], b1[3], a0[3], b0[3], v1[3], v0[3], v[3];
            FN(d2n_point_)(p, x, y, d, i00, i02, i11, i12);
            FN(d2n_point_)(p_y0, x, y-1, d_y0, i00, i02, i11, i12);
            FN(d2n_point_)(p_x0, x-1, y, d_x0, i00, i02, i11, i12);
            FN(d2n_point_)(p_y1, x, y+1, d_y1, i00, i02, i11, i12);
            FN(d2n_point_)(p_x1, x+1, y, d_x1, i00, i02, i11, i12);
            for (int k = 0; k < 3; ++k) { a1[k] = p[k] - p_x1[k]; b1[k] = p_y1[k] - p[k]; a0[k] = p[k] - p_x0[k]; b0[k] = p_y0[k] - p[k]; }
            FN(d2n_cross_)(v1, a1, b1);
            FN(d2n_cross_)(v0, a0, b0);
            FN(d2n_normalize_)(v1);
            FN(d2n_normalize_)(v0);
            for (int k = 0; k < 3; ++k) v[k] = v1[k] + v0[k];
            FN(d2n_normalize_)(v);
            n0 = v[0]; n1 = v[1]; n2 = v[2];
          }
        }
        normal[(long)y*x_size + x] = n0;
        normal[xy_size + (long)y*x_size + x] = n1;
        normal[2*xy_size + (long)y*x_size + x] = n2;
      }
  }
}
