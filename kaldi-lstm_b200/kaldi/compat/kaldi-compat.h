// compat/kaldi-compat.h -- TEST SCAFFOLDING, not a Kaldi re-implementation.
//
// The smallest slice of Kaldi nnet1's type surface that
// kaldi-lstm_b200/kaldi/b200-lstm-projected-streams.h needs, so that the component can be
// compiled and exercised in a tree that has no Kaldi checkout (none is vendored in the reference,
// SURVEY.md section 8c).  With a real Kaldi tree, define HAVE_KALDI and include the real headers
// instead; the component source is the same.
//
// Layout facts mirrored from the reference:
//   CuMatrixBase fields {data_, num_cols_, num_rows_, stride_}   google/cudamatrix/cu-matrix.h:479-489
//   pitched device allocation                                       google/cudamatrix/cu-matrix.cc:67-73
//   matrix / vector / token I/O formats                             google/matrix/kaldi-matrix.cc:1172-1211
#ifndef B200_KALDI_COMPAT_H_
#define B200_KALDI_COMPAT_H_
#include <cuda_runtime.h>
#include <stdint.h>

#include <cmath>
#include <cstring>
#include <iostream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace kaldi {
typedef float BaseFloat;
typedef int32_t int32;
typedef int32_t MatrixIndexT;
enum MatrixResizeType { kSetZero, kUndefined, kCopyData };
enum MatrixTransposeType { kTrans = 112, kNoTrans = 111 };

struct KaldiErr {
  std::ostringstream os;
  [[noreturn]] ~KaldiErr() noexcept(false) { throw std::runtime_error(os.str()); }
};
#define KALDI_ERR ::kaldi::KaldiErr().os
struct KaldiWarn {
  std::ostringstream os;
  ~KaldiWarn() { std::cerr << "WARNING " << os.str() << std::endl; }
};
#define KALDI_WARN ::kaldi::KaldiWarn().os
#define KALDI_ASSERT(cond)                                                            \
  do {                                                                                \
    if (!(cond)) throw std::runtime_error(std::string("KALDI_ASSERT failed: ") + #cond); \
  } while (0)
#define CU_SAFE_CALL(expr)                                                              \
  do {                                                                                  \
    cudaError_t e__ = (expr);                                                           \
    if (e__ != cudaSuccess) KALDI_ERR << "CUDA error: " << cudaGetErrorString(e__);      \
  } while (0)

// ---- token / basic-type I/O (base/io-funcs.h upstream) -------------------------------------
inline void WriteToken(std::ostream& os, bool, const std::string& t) { os << t << " "; }
inline void ReadToken(std::istream& is, bool binary, std::string* t) {
  if (!binary) is >> std::ws;
  is >> *t;
  if (is.fail()) KALDI_ERR << "ReadToken failed";
  is.get();  // the space after the token
}
inline void ExpectToken(std::istream& is, bool binary, const char* token) {
  std::string t;
  ReadToken(is, binary, &t);
  if (t != token) KALDI_ERR << "Expected token " << token << ", got " << t;
}
template <class T>
inline void WriteBasicType(std::ostream& os, bool binary, T v) {
  if (binary) {
    char sz = (char)sizeof(T);
    os.put(sz);
    os.write(reinterpret_cast<const char*>(&v), sizeof(T));
  } else {
    os << v << " ";
  }
}
template <class T>
inline void ReadBasicType(std::istream& is, bool binary, T* v) {
  if (binary) {
    int sz = is.get();
    if (sz != (int)sizeof(T)) KALDI_ERR << "ReadBasicType: size mismatch";
    is.read(reinterpret_cast<char*>(v), sizeof(T));
  } else {
    is >> *v;
  }
  if (is.fail()) KALDI_ERR << "ReadBasicType failed";
}

// ---- host matrix / vector --------------------------------------------------------------------
template <class Real>
class Vector {
 public:
  Vector() {}
  explicit Vector(MatrixIndexT n) : d_(n, Real(0)) {}
  void Resize(MatrixIndexT n) { d_.assign(n, Real(0)); }
  MatrixIndexT Dim() const { return (MatrixIndexT)d_.size(); }
  Real* Data() { return d_.data(); }
  const Real* Data() const { return d_.data(); }
  Real& operator()(MatrixIndexT i) { return d_[i]; }
  Real operator()(MatrixIndexT i) const { return d_[i]; }
  void Write(std::ostream& os, bool binary) const {
    if (binary) {
      WriteToken(os, true, "FV");
      WriteBasicType(os, true, (int32)Dim());
      os.write(reinterpret_cast<const char*>(Data()), sizeof(Real) * Dim());
    } else {
      os << " [ ";
      for (auto v : d_) os << v << " ";
      os << "]\n";
    }
  }
  void Read(std::istream& is, bool binary) {
    if (binary) {
      ExpectToken(is, true, "FV");
      int32 n;
      ReadBasicType(is, true, &n);
      Resize(n);
      is.read(reinterpret_cast<char*>(Data()), sizeof(Real) * n);
    } else {
      std::string t;
      is >> t;
      if (t != "[") KALDI_ERR << "vector: expected [";
      d_.clear();
      while (is >> t && t != "]") d_.push_back((Real)std::stod(t));
    }
  }

 private:
  std::vector<Real> d_;
};

// upstream Vector derives from VectorBase; the compat surface needs only the read accessors
template <class Real>
using VectorBase = Vector<Real>;
// hmm/posterior.h upstream
typedef std::vector<std::vector<std::pair<int32, BaseFloat> > > Posterior;

template <class Real>
class Matrix {
 public:
  Matrix() : r_(0), c_(0) {}
  Matrix(MatrixIndexT r, MatrixIndexT c) { Resize(r, c); }
  void Resize(MatrixIndexT r, MatrixIndexT c) {
    r_ = r;
    c_ = c;
    d_.assign((size_t)r * c, Real(0));
  }
  MatrixIndexT NumRows() const { return r_; }
  MatrixIndexT NumCols() const { return c_; }
  MatrixIndexT Stride() const { return c_; }
  Real* Data() { return d_.data(); }
  const Real* Data() const { return d_.data(); }
  Real& operator()(MatrixIndexT i, MatrixIndexT j) { return d_[(size_t)i * c_ + j]; }
  Real operator()(MatrixIndexT i, MatrixIndexT j) const { return d_[(size_t)i * c_ + j]; }
  void Write(std::ostream& os, bool binary) const {  // kaldi-matrix.cc:1172-1211
    if (binary) {
      WriteToken(os, true, "FM");
      WriteBasicType(os, true, (int32)r_);
      WriteBasicType(os, true, (int32)c_);
      os.write(reinterpret_cast<const char*>(Data()), sizeof(Real) * d_.size());
    } else {
      os << " [";
      for (MatrixIndexT i = 0; i < r_; i++) {
        os << "\n  ";
        for (MatrixIndexT j = 0; j < c_; j++) os << (*this)(i, j) << " ";
      }
      os << "]\n";
    }
  }
  void Read(std::istream& is, bool binary) {
    if (binary) {
      ExpectToken(is, true, "FM");
      int32 r, c;
      ReadBasicType(is, true, &r);
      ReadBasicType(is, true, &c);
      Resize(r, c);
      is.read(reinterpret_cast<char*>(Data()), sizeof(Real) * d_.size());
    } else {
      std::string line, t;
      is >> t;
      if (t != "[") KALDI_ERR << "matrix: expected [";
      std::vector<std::vector<Real> > rows;
      std::getline(is, line);  // rest of the "[" line
      bool done = false;
      while (!done && std::getline(is, line)) {
        std::istringstream ls(line);
        std::vector<Real> row;
        while (ls >> t) {
          if (t == "]") { done = true; break; }
          row.push_back((Real)std::stod(t));
        }
        if (!row.empty()) rows.push_back(row);
      }
      Resize((MatrixIndexT)rows.size(), rows.empty() ? 0 : (MatrixIndexT)rows[0].size());
      for (MatrixIndexT i = 0; i < r_; i++)
        for (MatrixIndexT j = 0; j < c_; j++) (*this)(i, j) = rows[i][j];
    }
  }

 private:
  MatrixIndexT r_, c_;
  std::vector<Real> d_;
};

// ---- device matrix: same field order as the reference (cu-matrix.h:479-489) ----------------------
template <class Real>
class CuMatrixBase {
 public:
  MatrixIndexT NumRows() const { return num_rows_; }
  MatrixIndexT NumCols() const { return num_cols_; }
  MatrixIndexT Stride() const { return stride_; }
  // In the reference snapshot Data()/RowData() are protected (cu-matrix.h:446-461); newer Kaldi makes
  // them public.  INTEGRATION.md lists the accessor options; the compat type simply exposes them.
  const Real* Data() const { return data_; }
  Real* Data() { return data_; }
  void SetZero() {
    if (data_) CU_SAFE_CALL(cudaMemset2D(data_, stride_ * sizeof(Real), 0, num_cols_ * sizeof(Real), num_rows_));
  }
  void CopyFromMat(const Matrix<Real>& m) {
    KALDI_ASSERT(m.NumRows() == num_rows_ && m.NumCols() == num_cols_);
    if (num_rows_)
      CU_SAFE_CALL(cudaMemcpy2D(data_, stride_ * sizeof(Real), m.Data(), m.Stride() * sizeof(Real),
                                num_cols_ * sizeof(Real), num_rows_, cudaMemcpyHostToDevice));
  }
  void CopyToMat(Matrix<Real>* m) const {
    m->Resize(num_rows_, num_cols_);
    if (num_rows_)
      CU_SAFE_CALL(cudaMemcpy2D(m->Data(), m->Stride() * sizeof(Real), data_, stride_ * sizeof(Real),
                                num_cols_ * sizeof(Real), num_rows_, cudaMemcpyDeviceToHost));
  }

 protected:
  CuMatrixBase() : data_(NULL), num_cols_(0), num_rows_(0), stride_(0) {}
  CuMatrixBase(Real* d, MatrixIndexT r, MatrixIndexT c, MatrixIndexT s) : data_(d), num_cols_(c), num_rows_(r), stride_(s) {}
  Real* data_;
  MatrixIndexT num_cols_;
  MatrixIndexT num_rows_;
  MatrixIndexT stride_;
};

template <class Real>
class CuMatrix : public CuMatrixBase<Real> {
 public:
  CuMatrix() {}
  CuMatrix(MatrixIndexT r, MatrixIndexT c, MatrixResizeType t = kSetZero) { Resize(r, c, t); }
  explicit CuMatrix(const Matrix<Real>& m) {
    Resize(m.NumRows(), m.NumCols(), kUndefined);
    this->CopyFromMat(m);
  }
  CuMatrix(const CuMatrix&) = delete;
  CuMatrix& operator=(const CuMatrix&) = delete;
  ~CuMatrix() { Destroy(); }
  void Resize(MatrixIndexT r, MatrixIndexT c, MatrixResizeType t = kSetZero) {  // cu-matrix.cc:50-84
    if (r == this->num_rows_ && c == this->num_cols_) {
      if (t == kSetZero) this->SetZero();
      return;
    }
    Destroy();
    if (r == 0 || c == 0) return;
    size_t pitch = 0;
    CU_SAFE_CALL(cudaMallocPitch((void**)&this->data_, &pitch, c * sizeof(Real), r));
    this->num_rows_ = r;
    this->num_cols_ = c;
    this->stride_ = (MatrixIndexT)(pitch / sizeof(Real));
    if (t == kSetZero) this->SetZero();
  }

 private:
  void Destroy() {
    if (this->data_) cudaFree(this->data_);
    this->data_ = NULL;
    this->num_rows_ = this->num_cols_ = this->stride_ = 0;
  }
};

namespace nnet1 {
struct NnetTrainOptions {  // nnet/nnet-trnopts.h upstream: the two fields the component reads
  BaseFloat learn_rate, momentum;
  NnetTrainOptions() : learn_rate(0.008f), momentum(0.0f) {}
};

class Component {  // nnet/nnet-component.h upstream (the part the LSTM component overrides / uses)
 public:
  enum ComponentType { kUnknown = 0x0, kLstmProjectedStreams = 0x0205 };
  Component(int32 in, int32 out) : input_dim_(in), output_dim_(out) {}
  virtual ~Component() {}
  virtual Component* Copy() const = 0;
  virtual ComponentType GetType() const = 0;
  virtual bool IsUpdatable() const { return false; }
  int32 InputDim() const { return input_dim_; }
  int32 OutputDim() const { return output_dim_; }
  // Component::Propagate / Backpropagate: size the output, then call the *Fnc hook
  void Propagate(const CuMatrixBase<BaseFloat>& in, CuMatrix<BaseFloat>* out) {
    KALDI_ASSERT(in.NumCols() == input_dim_);
    out->Resize(in.NumRows(), output_dim_, kSetZero);
    PropagateFnc(in, out);
  }
  void Backpropagate(const CuMatrixBase<BaseFloat>& in, const CuMatrixBase<BaseFloat>& out,
                     const CuMatrixBase<BaseFloat>& out_diff, CuMatrix<BaseFloat>* in_diff) {
    if (in_diff) in_diff->Resize(in.NumRows(), input_dim_, kSetZero);
    BackpropagateFnc(in, out, out_diff, in_diff);
  }
  virtual void Reset(std::vector<int>&) {}  // the no-op the reference's Nnet::Reset needs on every component
  virtual void InitData(std::istream&) {}
  virtual void ReadData(std::istream&, bool) {}
  virtual void WriteData(std::ostream&, bool) const {}

 protected:
  virtual void PropagateFnc(const CuMatrixBase<BaseFloat>& in, CuMatrixBase<BaseFloat>* out) = 0;
  virtual void BackpropagateFnc(const CuMatrixBase<BaseFloat>& in, const CuMatrixBase<BaseFloat>& out,
                                const CuMatrixBase<BaseFloat>& out_diff, CuMatrixBase<BaseFloat>* in_diff) = 0;
  int32 input_dim_, output_dim_;
};

class UpdatableComponent : public Component {
 public:
  UpdatableComponent(int32 in, int32 out) : Component(in, out) {}
  bool IsUpdatable() const { return true; }
  virtual int32 NumParams() const = 0;
  virtual void GetParams(Vector<BaseFloat>* v) const = 0;
  virtual void Update(const CuMatrixBase<BaseFloat>& input, const CuMatrixBase<BaseFloat>& diff) = 0;
  virtual void SetTrainOptions(const NnetTrainOptions& o) { opts_ = o; }
  const NnetTrainOptions& GetTrainOptions() const { return opts_; }

 protected:
  NnetTrainOptions opts_;
};
}  // namespace nnet1
}  // namespace kaldi
#endif
