// fimex_b200/Cached.h -- header-only C++ mirror of the reference's cached regridding operators, forwarding to the C ABI
// of libfimex_b200.so (include/fimex_b200.h).
//
// Reference classes (arebru/fimex 0.67.2):
//   MetNoFimex::CachedInterpolationInterface / CachedInterpolation   include/fimex/CachedInterpolation.h:60-161
//   MetNoFimex::CachedForwardInterpolation                           src/CachedForwardInterpolation.h:34-61
//   MetNoFimex::CachedVectorReprojection                             include/fimex/CachedVectorReprojection.h:33-63
// Same constructor arguments, same method names and meaning, same error behaviour (an exception where the reference
// throws CDMException).  boost::shared_array<float> becomes std::shared_ptr<float[]>; the reader-side methods
// (getInputDataSlice) stay in the host application because they only touch CDM metadata and I/O.
//
// In a Fimex tree the three .cc files would include this header and forward; stand-alone it lets
// the reference's own tests be restated nearly verbatim (tests/cpp/).
#ifndef FIMEX_B200_CACHED_H_
#define FIMEX_B200_CACHED_H_

#include "../fimex_b200.h"

#include <cstddef>
#include <cstring>
#include <list>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace MetNoFimexB200 {

class CDMException : public std::runtime_error
{
public:
    explicit CDMException(const std::string& msg) : std::runtime_error(msg) {}
};

typedef std::shared_ptr<float[]> shared_float_array;

struct ReducedInterpolationDomain
{
    std::string xDim, yDim;
    size_t xMin, yMin, xOrg, yOrg;
};

/** include/fimex/CachedInterpolation.h:60-99 */
class CachedInterpolationInterface
{
public:
    CachedInterpolationInterface(const std::string& xDimName, const std::string& yDimName)
        : _xDimName(xDimName), _yDimName(yDimName), handle_(0) {}
    virtual ~CachedInterpolationInterface() { fb200_interp_destroy(handle_); }

    /** interpolateValues(inData, size, newSize): [inZ][inY][inX] -> newly allocated [inZ][outY][outX] */
    virtual shared_float_array interpolateValues(shared_float_array inData, size_t size, size_t& newSize) const
    {
        newSize = fb200_interp_new_size(handle_, size);
        // page-locked and recycled by the library (fb200_host_alloc): the download runs at the full PCIe rate
        float* raw = static_cast<float*>(fb200_host_alloc(sizeof(float) * (newSize ? newSize : 1)));
        if (!raw)
            throw CDMException(std::string("error during interpolation: ") + fb200_last_error());
        shared_float_array out(raw, fb200_host_free);
        size_t n = 0;
        if (fb200_interp_interpolate_values(handle_, inData.get(), size, out.get(), &n) != MIFI_OK)
            throw CDMException(std::string("error during interpolation: ") + fb200_last_error());
        return out;
    }
    /** The whole per-slice body of CDMInterpolator::getDataSlice (src/CDMInterpolator.cc:250-258, 284-285) in one call:
     *  data2InterpolationArray(inData, badValue) -> interpolateValues -> interpolationArray2Data(type, ..., badValue).
     *  inType / outType: CDMDataType numbers (fb200_datatype); outData: outX*outY*inZ elements of outType, caller-owned.
     *  On the staged gathers both adapter passes run inside the gather kernel. */
    virtual void getDataSlice(int inType, const void* inData, size_t size, double badValue, int outType, void* outData, size_t& newSize) const
    {
        if (fb200_interp_get_data_slice(handle_, inType, inData, size, badValue, outType, outData, &newSize) != MIFI_OK)
            throw CDMException(std::string("error during interpolation: ") + fb200_last_error());
    }
    /** The vector branch of CDMInterpolator::getDataSlice (src/CDMInterpolator.cc:259-283) for BOTH components in one pass:
     *  fill -> NaN, interpolate u and v through one table, rotate (vector may be null), NaN -> fill + cast. */
    virtual void getVectorSlice(const fb200_vector* vector, int inType, const void* uIn, const void* vIn, size_t size, double badU, double badV,
                                int outType, void* uOut, void* vOut, size_t& newSize) const
    {
        if (fb200_interp_get_vector_slice(handle_, vector, inType, uIn, vIn, size, badU, badV, outType, uOut, vOut, &newSize) != MIFI_OK)
            throw CDMException(std::string("error during interpolation: ") + fb200_last_error());
    }
    virtual size_t getInX() const { return fb200_interp_in_x(handle_); }
    virtual size_t getInY() const { return fb200_interp_in_y(handle_); }
    virtual size_t getOutX() const { return fb200_interp_out_x(handle_); }
    virtual size_t getOutY() const { return fb200_interp_out_y(handle_); }
    std::shared_ptr<ReducedInterpolationDomain> reducedDomain() const { return reducedDomain_; }
    const fb200_interp* handle() const { return handle_; }

protected:
    std::string _xDimName, _yDimName;
    fb200_interp* handle_;
    std::shared_ptr<ReducedInterpolationDomain> reducedDomain_;

private:
    CachedInterpolationInterface(const CachedInterpolationInterface&);
    CachedInterpolationInterface& operator=(const CachedInterpolationInterface&);
};

/** include/fimex/CachedInterpolation.h:105-161, src/CachedInterpolation.cc:93-200 */
class CachedInterpolation : public CachedInterpolationInterface
{
public:
    CachedInterpolation(const std::string& xDimName, const std::string& yDimName, int funcType, const std::vector<double>& pointsOnXAxis,
                        const std::vector<double>& pointsOnYAxis, size_t inX, size_t inY, size_t outX, size_t outY)
        : CachedInterpolationInterface(xDimName, yDimName)
    {
        if (pointsOnXAxis.size() != outX * outY || pointsOnYAxis.size() != outX * outY)
            throw CDMException("pointsOnXAxis/pointsOnYAxis must hold outX*outY values");
        if (fb200_cached_interpolation_create(funcType, pointsOnXAxis.data(), pointsOnYAxis.data(), inX, inY, outX, outY, &handle_) != MIFI_OK)
            throw CDMException(std::string("unknown interpolation function or device error: ") + fb200_last_error());
    }

    /** src/CachedInterpolation.cc:159-200 */
    void createReducedDomain(std::string xDimName, std::string yDimName)
    {
        if (reducedDomain_)
            return; // don't set twice
        const size_t xOrg = getInX(), yOrg = getInY();
        int reduced = 0;
        long long x0 = 0, y0 = 0;
        if (fb200_interp_create_reduced_domain(handle_, &reduced, &x0, &y0) != MIFI_OK)
            throw CDMException(std::string("createReducedDomain: ") + fb200_last_error());
        if (!reduced)
            return;
        std::shared_ptr<ReducedInterpolationDomain> rid(new ReducedInterpolationDomain());
        rid->xDim = xDimName;
        rid->yDim = yDimName;
        rid->xMin = static_cast<size_t>(x0);
        rid->yMin = static_cast<size_t>(y0);
        rid->xOrg = xOrg;
        rid->yOrg = yOrg;
        reducedDomain_ = rid;
    }
};

/** src/CachedForwardInterpolation.h:34-61, src/CachedForwardInterpolation.cc:62-131 */
class CachedForwardInterpolation : public CachedInterpolationInterface
{
public:
    CachedForwardInterpolation(const std::string& xDimName, const std::string& yDimName, int funcType, const std::vector<double>& pOnX,
                               const std::vector<double>& pOnY, size_t inX, size_t inY, size_t outX, size_t outY)
        : CachedInterpolationInterface(xDimName, yDimName)
    {
        if (pOnX.size() != inX * inY || pOnY.size() != inX * inY)
            throw CDMException("pOnX/pOnY must hold inX*inY values");
        if (fb200_cached_forward_interpolation_create(funcType, pOnX.data(), pOnY.data(), inX, inY, outX, outY, &handle_) != MIFI_OK)
            throw CDMException(std::string("unknown forward interpolation method or device error: ") + fb200_last_error());
    }
};

/** include/fimex/CachedVectorReprojection.h:33-63, src/CachedVectorReprojection.cc:35-55 */
class CachedVectorReprojection
{
public:
    CachedVectorReprojection() : handle_(0), ox(0), oy(0) { fb200_vector_create(MIFI_VECTOR_KEEP_SIZE, 0, 0, 0, &handle_); }
    CachedVectorReprojection(int method, std::shared_ptr<double[]> matrix, int ox, int oy) : handle_(0), ox(ox), oy(oy)
    {
        if (fb200_vector_create(method, matrix.get(), ox, oy, &handle_) != MIFI_OK)
            throw CDMException(std::string("CachedVectorReprojection: ") + fb200_last_error());
    }
    /** makeCachedVectorReprojection(dataReader, cs, toLatLon) (src/CDMProcessor.cc:99-145): the matrix is built and kept on the
     *  device; axes in metres, or degrees when isDegree */
    CachedVectorReprojection(int method, const std::string& proj4, const std::vector<double>& xAxis, const std::vector<double>& yAxis,
                             bool isDegree, bool toLatLon)
        : handle_(0), ox(static_cast<int>(xAxis.size())), oy(static_cast<int>(yAxis.size()))
    {
        if (fb200_vector_create_from_grid(method, proj4.c_str(), xAxis.data(), yAxis.data(), ox, oy, isDegree ? 1 : 0, toLatLon ? 1 : 0,
                                          &handle_) != MIFI_OK)
            throw CDMException(std::string("makeCachedVectorReprojection: ") + fb200_last_error());
    }
    virtual ~CachedVectorReprojection() { fb200_vector_destroy(handle_); }

    void reprojectValues(shared_float_array& uValues, shared_float_array& vValues, size_t size) const
    {
        if (fb200_vector_reproject_values(handle_, uValues.get(), vValues.get(), size) != MIFI_OK)
            throw CDMException("Error during reprojection of vector-values");
    }
    void reprojectDirectionValues(shared_float_array& angles, size_t size) const
    {
        if (fb200_vector_reproject_direction_values(handle_, angles.get(), size) != MIFI_OK)
            throw CDMException("Error during reprojection of vector-direction-values");
    }
    /** the rotation branch of CDMProcessor::getDataSlice (src/CDMProcessor.cc:579-617) on raw typed slices */
    void getVectorSlice(fb200_datatype inType, const void* uIn, const void* vIn, size_t size, double badU, double badV, fb200_datatype outType,
                        void* uOut, void* vOut) const
    {
        if (fb200_vector_get_slice(handle_, inType, uIn, vIn, size, badU, badV, outType, uOut, vOut) != MIFI_OK)
            throw CDMException(std::string("Error during reprojection of vector-values: ") + fb200_last_error());
    }
    size_t getXSize() const { return static_cast<size_t>(ox); }
    size_t getYSize() const { return static_cast<size_t>(oy); }
    const fb200_vector* handle() const { return handle_; }

private:
    fb200_vector* handle_;
    int ox, oy;
    CachedVectorReprojection(const CachedVectorReprojection&);
    CachedVectorReprojection& operator=(const CachedVectorReprojection&);
};

/** The other half of an x/y vector pair, parked between the two getDataSlice calls of a pair.
 *
 *  The reference interpolates and rotates BOTH components on each component's call and throws one of them away
 *  (src/CDMInterpolator.cc:259-276): 2x redundant work.  With this cache the first call of a pair computes both halves with
 *  CachedInterpolationInterface::getVectorSlice, returns its own and parks the counterpart under a key that names the pair and
 *  the slice (e.g. "x_wind|y_wind|<unLimDimPos>"); the counterpart's call takes it out instead of computing again.
 *
 *  A parked half is handed out once.  Bounded (oldest entries are dropped), so a host that never asks for the counterpart only
 *  loses the saving.  Thread-safe: getDataSlice is called concurrently from OpenMP tasks (src/NetCDF_CDMWriter.cc:749-753).
 *  The host drops the cache when the interpolation changes (CDMInterpolator::changeProjection). */
class VectorPairCache
{
public:
    typedef std::shared_ptr<unsigned char[]> bytes;
    explicit VectorPairCache(size_t slots = 8) : slots_(slots ? slots : 1) {}

    /** park `size` bytes of the counterpart (direction 0 = x, 1 = y) of the pair `key` */
    void park(const std::string& key, int direction, const void* data, size_t size)
    {
        bytes copy(new unsigned char[size ? size : 1]);
        std::memcpy(copy.get(), data, size);
        std::lock_guard<std::mutex> lock(mu_);
        while (entries_.size() >= slots_)
            entries_.pop_front(); // oldest first
        entries_.push_back(Entry{key, direction, size, copy});
    }
    /** take the parked half out (true and `size` bytes copied to `out`), or false if it is not there (any more) */
    bool take(const std::string& key, int direction, void* out, size_t size)
    {
        bytes found;
        {
            std::lock_guard<std::mutex> lock(mu_);
            for (std::list<Entry>::iterator it = entries_.begin(); it != entries_.end(); ++it) {
                if (it->direction == direction && it->size == size && it->key == key) {
                    found = it->data;
                    entries_.erase(it);
                    break;
                }
            }
        }
        if (!found)
            return false;
        std::memcpy(out, found.get(), size);
        return true;
    }
    void clear()
    {
        std::lock_guard<std::mutex> lock(mu_);
        entries_.clear();
    }
    size_t size() const
    {
        std::lock_guard<std::mutex> lock(mu_);
        return entries_.size();
    }

    /** One component of an x/y pair as CDMInterpolator::getDataSlice returns it, both halves computed at most once:
     *  `direction` 0: `data` is the x component and `counterpart` the y component, 1: the other way round. */
    void getDataSlice(const CachedInterpolationInterface& ci, const fb200_vector* vector, const std::string& key, int direction, int inType,
                      const void* data, const void* counterpart, size_t size, double bad, double badCounterpart, int outType, size_t outElemSize,
                      void* out, size_t& newSize)
    {
        newSize = fb200_interp_new_size(ci.handle(), size);
        const size_t outBytes = newSize * outElemSize;
        if (take(key, direction, out, outBytes))
            return;
        bytes other(new unsigned char[outBytes ? outBytes : 1]);
        if (direction == 0)
            ci.getVectorSlice(vector, inType, data, counterpart, size, bad, badCounterpart, outType, out, other.get(), newSize);
        else
            ci.getVectorSlice(vector, inType, counterpart, data, size, badCounterpart, bad, outType, other.get(), out, newSize);
        std::lock_guard<std::mutex> lock(mu_);
        while (entries_.size() >= slots_)
            entries_.pop_front();
        entries_.push_back(Entry{key, 1 - direction, outBytes, other});
    }

private:
    struct Entry {
        std::string key;
        int direction;
        size_t size;
        bytes data;
    };
    size_t slots_;
    mutable std::mutex mu_;
    std::list<Entry> entries_;
};

} // namespace MetNoFimexB200

#endif
