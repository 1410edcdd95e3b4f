// apbf_pbd.hpp -- the reference's operator surface (namespace pbd) on top of the C-ABI of libapbf_b200.so.
//
// Header-only C++17.  A caller of cg-tuwien/APBF's particle hot path keeps its code: the list abstractions
//   pbd::gpu_list<Stride>          (source/gpu_list.h:11-57)
//   pbd::indexed_list<DataList>    (source/indexed_list.h:13-79)
//   pbd::uninterleaved_list<E,...> (source/uninterleaved_list.h:7-83)
//   pbd::hidden_particles / particles / fluid / neighbors (source/list_definitions.h:9-20)
// and the operators
//   pbd::algorithms                (source/algorithms.h:12-17)
//   pbd::neighborhood_green        (source/neighborhood_green.h:11-14)
//   pbd::neighborhood_binary_search(source/neighborhood_binary_search.h:11-13)
//   pbd::incompressibility         (source/incompressibility.h:11-12)
//   pbd::spread_kernel_width       (source/spread_kernel_width.h:10-11)
//   pbd::box_collision             (source/box_collision.h:11-12)
//   pbd::velocity_handling         (source/velocity_handling.h:11-13)
// have the reference's names, signatures, ownership rules (lazy copy, write() makes unique, operators keep raw
// non-owning pointers) and ordering (one ordered stream of work).  What changes: avk::buffer becomes
// pbd::buffer (a device pointer + byte size), glm::vec3 becomes pbd::vec3, and shader_provider::set_queue /
// start_recording / end_recording become pbd::shader_provider::set_context / start_recording / end_recording over an
// apbf_ctx.  Errors of the C-ABI surface as std::runtime_error (the reference asserts); there is no CPU fallback.
#pragma once

#include <algorithm>
#include <cassert>
#include <cmath>
#include <filesystem>
#include <fstream>
#include <cstdint>
#include <cstring>
#include <limits>
#include <list>
#include <memory>
#include <stdexcept>
#include <string>
#include <tuple>
#include <utility>
#include <vector>

#include "apbf_b200.h"

namespace pbd
{
	struct vec3 { float x = 0.f, y = 0.f, z = 0.f; vec3() = default; vec3(float a, float b, float c) : x(a), y(b), z(c) {} explicit vec3(float a) : x(a), y(a), z(a) {} };
	inline vec3 operator-(const vec3& a, float b) { return vec3(a.x - b, a.y - b, a.z - b); }
	inline vec3 operator+(const vec3& a, float b) { return vec3(a.x + b, a.y + b, a.z + b); }

	/// what an avk::buffer is to the reference: a device allocation (never owned by this struct)
	struct buffer
	{
		void*  ptr = nullptr;
		size_t bytes = 0;
		template<class T> T* as() const { return static_cast<T*>(ptr); }
	};

	/// shader_provider's recording state (source/shader_provider.cpp:5-37) reduced to "the current context"
	class shader_provider
	{
	public:
		static void set_context(apbf_ctx* aCtx) { current() = aCtx; }
		static apbf_ctx* context()
		{
			if (current() == nullptr) throw std::runtime_error("pbd::shader_provider::set_context() has not been called");
			return current();
		}
		static void start_recording() { recording() = true; }                        // work is enqueued immediately on the context's stream
		static void end_recording() { recording() = false; check(apbf_ctx_synchronize(context())); } // the reference submits and waits here
		static bool is_recording() { return recording(); }
		static void check(int aStatus)
		{
			if (aStatus != APBF_OK) throw std::runtime_error(std::string("apbf_b200: ") + apbf_ctx_last_error(current()) + " (status " + std::to_string(aStatus) + ")");
		}
		// the list helper shaders (shader_provider.h "gpu_lists" group); lengths are device words
		static void write_sequence(const buffer& aOut, const buffer& aLength, uint32_t aStart, uint32_t aStep)
		{
			check(apbf_write_sequence(context(), aOut.as<uint32_t>(), aLength.as<uint32_t>(), static_cast<uint32_t>(aOut.bytes / 4), aStart, aStep, 1u));
		}
		/// shader_provider.h:27 (pool.cpp:77-80 calls it with the boundary distance, the kernel width, the particle index list and the radii)
		static void uint_to_float_with_indexed_lower_bound(const buffer& aInUintBuffer, const buffer& aOutFloatBuffer, const buffer& aInIndexList, const buffer& aInLowerBound,
		                                                   const buffer& aInUintBufferLength, float aFactor, float aLowerBoundFactor, float aMaxAdationStep)
		{
			check(apbf_uint_to_float_with_indexed_lower_bound(context(), aInUintBuffer.as<uint32_t>(), aOutFloatBuffer.as<float>(), aInIndexList.as<uint32_t>(), aInLowerBound.as<float>(),
			                                                  aInUintBufferLength.as<uint32_t>(), static_cast<uint32_t>(std::min(aInUintBuffer.bytes, aOutFloatBuffer.bytes) / 4), aFactor,
			                                                  aLowerBoundFactor, aMaxAdationStep));
		}
		static void write_sequence_float(const buffer& aOut, const buffer& aLength, float aStart, float aStep)
		{
			check(apbf_write_sequence_float(context(), aOut.as<float>(), aLength.as<uint32_t>(), static_cast<uint32_t>(aOut.bytes / 4), aStart, aStep));
		}
	private:
		static apbf_ctx*& current() { static apbf_ctx* c = nullptr; return c; }
		static bool& recording() { static bool r = false; return r; }
	};

	/// storage of one list: data buffer + 4-byte length word, both from the context's best-fit pool
	/// (source/gpu_list_data.{h,cpp})
	class gpu_list_data
	{
	public:
		buffer mBuffer;
		buffer mLength;
		static std::shared_ptr<gpu_list_data> get_list(size_t aMinLength, size_t aStride)
		{
			auto d = std::shared_ptr<gpu_list_data>(new gpu_list_data());
			d->mCtx = shader_provider::context();
			d->mBuffer.bytes = std::max<size_t>(aMinLength * aStride, 4);
			shader_provider::check(apbf_buffer_acquire(d->mCtx, d->mBuffer.bytes, &d->mBuffer.ptr));
			d->mLength.bytes = 4;
			shader_provider::check(apbf_buffer_acquire(d->mCtx, 4, &d->mLength.ptr));
			return d;
		}
		~gpu_list_data()
		{
			if (mBuffer.ptr) apbf_buffer_release(mCtx, mBuffer.ptr);
			if (mLength.ptr) apbf_buffer_release(mCtx, mLength.ptr);
		}
		gpu_list_data(const gpu_list_data&) = delete;
		gpu_list_data& operator=(const gpu_list_data&) = delete;
	private:
		gpu_list_data() = default;
		apbf_ctx* mCtx = nullptr;
	};

	class algorithms
	{
	public:
		algorithms() = delete;
		static void copy_bytes(const buffer& aSource, const buffer& aTarget, size_t aCopiedLength, size_t aSourceOffset = 0, size_t aTargetOffset = 0)
		{
			assert(aSourceOffset + aCopiedLength <= aSource.bytes && aTargetOffset + aCopiedLength <= aTarget.bytes);
			shader_provider::check(apbf_copy_bytes(shader_provider::context(), static_cast<const char*>(aSource.ptr) + aSourceOffset, static_cast<char*>(aTarget.ptr) + aTargetOffset, aCopiedLength));
		}
		static void copy_bytes(const void* aSource, const buffer& aTarget, size_t aCopiedLength, size_t aSourceOffset = 0, size_t aTargetOffset = 0)
		{
			assert(aTargetOffset + aCopiedLength <= aTarget.bytes);
			// the host memory may be a temporary: the copy has completed when this returns
			shader_provider::check(apbf_copy_bytes_from_host(shader_provider::context(), static_cast<const char*>(aSource) + aSourceOffset, static_cast<char*>(aTarget.ptr) + aTargetOffset, aCopiedLength));
			shader_provider::check(apbf_ctx_synchronize(shader_provider::context()));
		}
		static size_t sort_calculate_needed_helper_list_length(size_t aMaxValueCount) { return apbf_sort_calculate_needed_helper_list_length(aMaxValueCount); }
		static size_t prefix_sum_calculate_needed_helper_list_length(size_t aMaxValueCount) { return apbf_prefix_sum_calculate_needed_helper_list_length(aMaxValueCount); }
		/// stable, ascending, key + payload; the helper list is accepted for source compatibility and not used
		static void sort(const buffer& aValues, const buffer& aSecondValues, const buffer& /*aHelperList*/, const buffer& aValueCount, size_t aMaxValueCount, const buffer& aResult, const buffer& aSecondResult, uint32_t aValueUpperBound = std::numeric_limits<uint32_t>::max())
		{
			shader_provider::check(apbf_sort(shader_provider::context(), aValues.as<uint32_t>(), aSecondValues.as<uint32_t>(), aValueCount.as<uint32_t>(), static_cast<uint32_t>(aMaxValueCount), aResult.as<uint32_t>(), aSecondResult.as<uint32_t>(), aValueUpperBound));
		}
		/// inclusive; in place unless aResult is given
		static void prefix_sum(const buffer& aValues, const buffer& /*aHelperList*/, const buffer& aValueCount, size_t aMaxValueCount, const buffer* aResult = nullptr)
		{
			shader_provider::check(apbf_prefix_sum(shader_provider::context(), aValues.as<uint32_t>(), aValueCount.as<uint32_t>(), static_cast<uint32_t>(aMaxValueCount), (aResult ? *aResult : aValues).as<uint32_t>()));
		}
	};

	template<class EditList>
	class list_interface
	{
	public:
		virtual ~list_interface() = default;
		virtual void apply_edit(EditList& pEditList, list_interface* pEditSource) = 0;
	};

	/// A list that lives in device memory only.  Copies share the storage until one of them asks for write().
	template<size_t Stride>
	class gpu_list : public list_interface<gpu_list<4>>
	{
	public:
		using edit_list = gpu_list<4>;
		using owner_ptr = list_interface<gpu_list<4>>*;

		gpu_list() = default;
		gpu_list(const gpu_list& aOther) : mData(aOther.mData), mRequestedLength(aOther.mRequestedLength) {} // the owner is not copied
		~gpu_list() override = default;

		/// "no device buffer assigned yet", not "zero elements" (the host does not know the length)
		bool empty() const { return mData == nullptr; }
		pbd::buffer& length() const { assert(mData); return mData->mLength; }
		gpu_list& set_length(size_t aLength)
		{
			mRequestedLength = std::max(mRequestedLength, aLength);
			if (aLength == 0 && mData && mData.use_count() > 1) mData = nullptr; // do not copy content that is about to be dropped
			const uint32_t l = static_cast<uint32_t>(aLength);
			if (write().mData) algorithms::copy_bytes(&l, mData->mLength, 4);
			return *this;
		}
		gpu_list& set_length(const pbd::buffer& aLength)
		{
			if (write().mData && mData->mLength.ptr != aLength.ptr) algorithms::copy_bytes(aLength, mData->mLength, 4);
			return *this;
		}
		gpu_list& request_length(size_t aLength) { mRequestedLength = aLength; return *this; }
		size_t requested_length() const { return mRequestedLength; }

		void apply_edit(edit_list& aEditList, owner_ptr aEditSource) override
		{
			if (mOwner != aEditSource && mOwner != nullptr) mOwner->apply_edit(aEditList, this);
			if (static_cast<owner_ptr>(this) == aEditSource) return;
			auto old = mData;
			if (!old) return;
			fresh_storage(std::max<size_t>(mRequestedLength, aEditList.capacity()), false);
			shader_provider::check(apbf_copy_scattered_read(shader_provider::context(), old->mBuffer.ptr, mData->mBuffer.ptr, aEditList.buffer().template as<uint32_t>(),
			                                                aEditList.length().template as<uint32_t>(), std::min(capacity(), aEditList.capacity()), static_cast<uint32_t>(Stride)));
			algorithms::copy_bytes(aEditList.length(), mData->mLength, 4);
		}

		gpu_list& operator=(const gpu_list& aRhs) { mData = aRhs.mData; mRequestedLength = aRhs.mRequestedLength; return *this; }
		gpu_list& operator+=(const gpu_list& aRhs)
		{
			if (!mData) { mData = aRhs.mData; return *this; }
			if (!aRhs.mData) return *this;
			write();
			shader_provider::check(apbf_append_list(shader_provider::context(), mData->mBuffer.ptr, aRhs.mData->mBuffer.ptr, mData->mLength.template as<uint32_t>(),
			                                        aRhs.mData->mLength.template as<uint32_t>(), mData->mLength.template as<uint32_t>(), capacity(), aRhs.capacity(), static_cast<uint32_t>(Stride)));
			return *this;
		}
		gpu_list operator+(const gpu_list& aRhs) const
		{
			gpu_list r = *this;
			r.mRequestedLength = std::max(mRequestedLength, aRhs.mRequestedLength);
			r += aRhs;
			return r;
		}

		gpu_list& set_owner(owner_ptr aOwner) { mOwner = aOwner; return *this; }
		owner_ptr owner() const { return mOwner; }
		/// valid until another member of this object is called or the object is copied
		pbd::buffer& buffer() const { assert(mData != nullptr); return mData->mBuffer; }
		/// host read-back, debugging only (synchronises)
		template<class T>
		std::vector<T> read(bool aBeyondLength = false) const
		{
			std::vector<T> data;
			if (empty()) return data;
			data.resize(mData->mBuffer.bytes / sizeof(T));
			shader_provider::check(apbf_copy_bytes_to_host(shader_provider::context(), mData->mBuffer.ptr, data.data(), data.size() * sizeof(T)));
			if (!aBeyondLength) {
				uint32_t len = 0;
				shader_provider::check(apbf_copy_bytes_to_host(shader_provider::context(), mData->mLength.ptr, &len, 4));
				data.resize(std::min<size_t>(data.size(), static_cast<size_t>(len) * Stride / sizeof(T)));
			}
			return data;
		}
		/// write access: makes the storage unique and at least requested_length() long, keeping the content.  For a pair list
		/// (Stride 8) the caller may be about to rewrite pairs with kernels of its own: whatever the operators remember about
		/// this buffer is out of date from here on (they rebuild it from the pairs at their next apply()).
		gpu_list& write()
		{
			make_unique();
			if (Stride == 8 && mData) { apbf_neighbors nb{ mData->mBuffer.template as<uint32_t>(), mData->mLength.template as<uint32_t>(), capacity() }; apbf_neighbors_invalidate(shader_provider::context(), &nb); }
			return *this;
		}
		/// what write() does to the storage, without telling the operators that the content changes (used by the operators themselves)
		gpu_list& make_unique() { if (mRequestedLength != 0 || mData) { if (!unique_and_large_enough(mRequestedLength)) fresh_storage(mRequestedLength, true); } return *this; }

		/// gpu_list<4> only: interprets the values as uint and sorts them ascending; the owner follows the permutation
		void sort(size_t aValueUpperBound = std::numeric_limits<uint32_t>::max())
		{
			static_assert(Stride == 4, "sort() is defined for gpu_list<4>");
			if (!mData) return;
			gpu_list<4> unsorted = *this; // shares the storage: this->write() below moves *this to a fresh buffer
			auto unsortedIndices = gpu_list<4>().request_length(requested_length());
			auto sortedIndices = gpu_list<4>().request_length(requested_length());
			sortedIndices.set_length(length());
			unsortedIndices.set_length(length());
			shader_provider::write_sequence(unsortedIndices.write().buffer(), length(), 0u, 1u);
			fresh_storage(std::max<size_t>(mRequestedLength, capacity()), false);
			algorithms::copy_bytes(unsorted.length(), mData->mLength, 4);
			algorithms::sort(unsorted.write().buffer(), unsortedIndices.write().buffer(), pbd::buffer(), unsorted.length(), unsorted.capacity(), mData->mBuffer, sortedIndices.write().buffer(), static_cast<uint32_t>(aValueUpperBound));
			if (mOwner != nullptr) mOwner->apply_edit(sortedIndices, this);
		}

		template<size_t NewStride>
		gpu_list<NewStride> convert_to_stride() const
		{
			auto result = gpu_list<NewStride>().request_length(mRequestedLength);
			if (!mData) return result;
			result.set_length(length());
			shader_provider::check(apbf_copy_with_differing_stride(shader_provider::context(), mData->mBuffer.ptr, result.write().buffer().ptr, mData->mLength.template as<uint32_t>(),
			                                                       std::min(capacity(), result.capacity()), static_cast<uint32_t>(Stride), static_cast<uint32_t>(NewStride)));
			return result;
		}

		// ---- extensions used by the fused operators below (not part of the reference interface) ----
		uint32_t capacity() const { return mData ? static_cast<uint32_t>(mData->mBuffer.bytes / Stride) : 0u; }
		/// Moves the list to fresh storage of requested_length() elements and returns the previous storage (kept alive
		/// by the caller while the device reads it): the copy-on-write target of gpu_list::apply_edit (gpu_list.h:131-133).
		std::shared_ptr<gpu_list_data> rewrite_begin()
		{
			auto old = mData;
			fresh_storage(std::max<size_t>(mRequestedLength, old ? old->mBuffer.bytes / Stride : 0), false);
			if (old) algorithms::copy_bytes(old->mLength, mData->mLength, 4);
			return old;
		}

	private:
		template<size_t> friend class gpu_list;
		bool unique_and_large_enough(size_t aNeeded) const { return mData && mData.use_count() == 1 && mData->mBuffer.bytes / Stride >= aNeeded; }
		/// new storage; the length word (and, if asked for, the content) is carried over
		std::shared_ptr<gpu_list_data> fresh_storage(size_t aNeeded, bool aKeepContent)
		{
			auto old = mData;
			if (aNeeded == 0 && !old) return old;
			if (aNeeded == 0) { mData = nullptr; return old; }
			mData = gpu_list_data::get_list(aNeeded, Stride);
			if (!old) {
				const uint32_t zero = 0u;
				algorithms::copy_bytes(&zero, mData->mLength, 4);
			} else {
				algorithms::copy_bytes(old->mLength, mData->mLength, 4);
				if (aKeepContent) algorithms::copy_bytes(old->mBuffer, mData->mBuffer, std::min(old->mBuffer.bytes, aNeeded * Stride));
			}
			return old;
		}

		std::shared_ptr<gpu_list_data> mData;
		size_t mRequestedLength = 0;
		owner_ptr mOwner = nullptr;
	};

	/// A tuple of lists of equal length addressed by an enum; length changes and edits reach every member.
	template<class NameEnum, class... Lists>
	class uninterleaved_list : public list_interface<gpu_list<4>>
	{
	public:
		using id = NameEnum;
		using owner_ptr = list_interface<gpu_list<4>>*;

		uninterleaved_list() { adopt(); }
		uninterleaved_list(const uninterleaved_list& aOther) : mLists(aOther.mLists) { adopt(); }
		~uninterleaved_list() override = default;

		bool empty() const { return std::get<0>(mLists).empty(); }
		buffer& length() const { return std::get<0>(mLists).length(); }
		uninterleaved_list& set_length(size_t aLength) { std::apply([&](auto&... l) { (l.set_length(aLength), ...); }, mLists); return *this; }
		uninterleaved_list& set_length(const buffer& aLength) { std::apply([&](auto&... l) { (l.set_length(aLength), ...); }, mLists); return *this; }
		uninterleaved_list& request_length(size_t aLength) { std::apply([&](auto&... l) { (l.request_length(aLength), ...); }, mLists); return *this; }
		size_t requested_length() { return std::get<0>(mLists).requested_length(); }
		void apply_edit(gpu_list<4>& aEditList, owner_ptr aEditSource) override
		{
			if (mOwner != aEditSource && mOwner != nullptr) mOwner->apply_edit(aEditList, this);
			std::apply([&](auto&... l) { ((static_cast<owner_ptr>(&l) == aEditSource ? void() : l.apply_edit(aEditList, this)), ...); }, mLists);
		}
		uninterleaved_list& operator=(const uninterleaved_list& aRhs) { mLists = aRhs.mLists; adopt(); return *this; }
		uninterleaved_list& operator+=(const uninterleaved_list& aRhs) { add(aRhs, std::index_sequence_for<Lists...>()); return *this; }
		uninterleaved_list operator+(const uninterleaved_list& aRhs) const { auto r = *this; r += aRhs; return r; }
		uninterleaved_list& write() { std::apply([&](auto&... l) { (l.write(), ...); }, mLists); return *this; }
		uninterleaved_list& set_owner(owner_ptr aOwner) { mOwner = aOwner; return *this; }
		owner_ptr owner() const { return mOwner; }
		template<NameEnum E> constexpr auto& get() { return std::get<static_cast<int>(E)>(mLists); }
		template<NameEnum E> constexpr auto& get() const { return std::get<static_cast<int>(E)>(mLists); }

	private:
		void adopt() { std::apply([&](auto&... l) { (l.set_owner(this), ...); }, mLists); }
		template<size_t... Is> void add(const uninterleaved_list& aRhs, std::index_sequence<Is...>) { ((std::get<Is>(mLists) += std::get<Is>(aRhs.mLists)), ...); }
		std::tuple<Lists...> mLists;
		owner_ptr mOwner = nullptr;
	};

	/// Index list + shared hidden list.  List manipulation targets the index list; edits of the hidden list re-map
	/// the index list of every indexed_list that shares it (and, through set_owner, the lists that run parallel to it).
	template<class DataList>
	class indexed_list : public list_interface<gpu_list<4>>
	{
	public:
		using owner_ptr = list_interface<gpu_list<4>>*;

		indexed_list(size_t aAllocatedHiddenDataLength = 0) : mHiddenData(std::make_shared<hidden_data>())
		{
			mHiddenData->mData.request_length(aAllocatedHiddenDataLength);
			mIndexList.set_owner(this);
			mHiddenData->mOwners.push_back(this);
		}
		indexed_list(const indexed_list& aOther) : mIndexList(aOther.mIndexList), mHiddenData(aOther.mHiddenData), mSorted(aOther.mSorted)
		{
			mIndexList.set_owner(this);
			mHiddenData->mOwners.push_back(this);
		}
		~indexed_list() override { mHiddenData->mOwners.remove(this); }

		indexed_list& share_hidden_data_from(const indexed_list& aBenefactor)
		{
			mHiddenData->mOwners.remove(this);
			mHiddenData = aBenefactor.mHiddenData;
			mHiddenData->mOwners.push_back(this);
			return *this;
		}
		/// removes the hidden entries this list points to (from every list that shares them)
		void delete_these()
		{
			auto& hidden = mHiddenData->mData;
			const auto helperLength = hidden.requested_length();
			auto keep = gpu_list<4>().request_length(helperLength);
			keep.set_length(hidden.length());
			shader_provider::write_sequence(keep.write().buffer(), hidden.length(), 1u, 0u);
			shader_provider::check(apbf_scattered_write(shader_provider::context(), mIndexList.buffer().template as<uint32_t>(), keep.write().buffer().template as<uint32_t>(),
			                                            mIndexList.length().template as<uint32_t>(), mIndexList.capacity(), 0u));
			mIndexList.set_length(0);
			algorithms::prefix_sum(keep.write().buffer(), buffer(), hidden.length(), helperLength);
			auto editList = gpu_list<4>().request_length(helperLength);
			editList.set_length(0);
			shader_provider::check(apbf_find_value_changes(shader_provider::context(), keep.buffer().template as<uint32_t>(), editList.write().buffer().template as<uint32_t>(),
			                                               hidden.length().template as<uint32_t>(), editList.length().template as<uint32_t>(), static_cast<uint32_t>(helperLength)));
			mHiddenData->apply_edit(editList, this);
		}
		/// appends copies of the hidden entries this list points to; returns a list of the copies
		indexed_list duplicate_these()
		{
			auto& hidden = mHiddenData->mData;
			auto editList = gpu_list<4>().request_length(hidden.requested_length());
			editList.set_length(hidden.length());
			auto newIndices = gpu_list<4>().request_length(mIndexList.requested_length());
			newIndices.set_length(0);
			shader_provider::write_sequence(editList.write().buffer(), hidden.length(), 0u, 1u);
			editList += mIndexList;
			shader_provider::check(apbf_write_increasing_sequence_from_to(shader_provider::context(), newIndices.write().buffer().template as<uint32_t>(), newIndices.length().template as<uint32_t>(),
			                                                              hidden.length().template as<uint32_t>(), editList.length().template as<uint32_t>(), newIndices.capacity()));
			mHiddenData->apply_edit(editList, this);
			indexed_list result = *this;
			result.mIndexList = newIndices;
			result.mIndexList.set_owner(&result);
			return result;
		}

		bool empty() const { return mIndexList.empty(); }
		buffer& length() const { return mIndexList.length(); }
		indexed_list& set_length(size_t aLength) { mIndexList.set_length(aLength); return *this; }
		indexed_list& set_length(const buffer& aLength) { mIndexList.set_length(aLength); return *this; }
		indexed_list& request_length(size_t aLength) { mIndexList.request_length(aLength); return *this; }
		size_t requested_length() { return mIndexList.requested_length(); }
		/// appends aAddedLength indices that point to so far unused hidden entries; returns a list of just the new ones
		indexed_list increase_length(size_t aAddedLength)
		{
			indexed_list result = indexed_list().share_hidden_data_from(*this).request_length(mIndexList.requested_length());
			if (aAddedLength == 0) { result.write(); return result; }
			auto& hidden = mHiddenData->mData;
			hidden.write();
			if (hidden.empty()) throw std::runtime_error("indexed_list::increase_length(): the hidden list has no requested length");
			result.write();
			auto newHiddenLength = gpu_list<4>().request_length(1);
			newHiddenLength.write();
			shader_provider::check(apbf_write_increasing_sequence(shader_provider::context(), result.index_buffer().template as<uint32_t>(), result.mIndexList.capacity(),
			                                                      result.length().template as<uint32_t>(), hidden.length().template as<uint32_t>(), newHiddenLength.buffer().template as<uint32_t>(),
			                                                      static_cast<uint32_t>(hidden.requested_length()), static_cast<uint32_t>(aAddedLength)));
			hidden.set_length(newHiddenLength.buffer()); // every member list of an uninterleaved hidden list has its own length word
			result.mSorted = true;
			*this += result;
			return result;
		}
		void apply_edit(gpu_list<4>& aEditList, owner_ptr aEditSource) override
		{
			mSorted = false;
			if (static_cast<owner_ptr>(&mIndexList) != aEditSource) mIndexList.apply_edit(aEditList, this);
			if (mOwner != aEditSource && mOwner != nullptr) mOwner->apply_edit(aEditList, this);
		}
		indexed_list& operator=(const indexed_list& aRhs)
		{
			if (this == &aRhs) return *this;
			mHiddenData->mOwners.remove(this);
			mIndexList = aRhs.mIndexList;
			mHiddenData = aRhs.mHiddenData;
			mIndexList.set_owner(this);
			mHiddenData->mOwners.push_back(this);
			mSorted = aRhs.mSorted;
			return *this;
		}
		indexed_list& operator+=(const indexed_list& aRhs)
		{
			if (mHiddenData->mData.empty()) share_hidden_data_from(aRhs);
			mIndexList += aRhs.mIndexList;
			mSorted = false;
			return *this;
		}
		indexed_list operator+(const indexed_list& aRhs) const { auto r = *this; r += aRhs; return r; }
		void set_owner(owner_ptr aOwner) { mOwner = aOwner; }
		owner_ptr owner() const { return mOwner; }
		buffer& index_buffer() const { return mIndexList.buffer(); }
		DataList& hidden_list() { return mHiddenData->mData; }
		std::vector<uint32_t> index_read(bool aBeyondLength = false) const { return mIndexList.template read<uint32_t>(aBeyondLength); }
		indexed_list& write() { mIndexList.write(); mSorted = false; return *this; }
		void sort(size_t aValueUpperBound) { if (!mSorted) mIndexList.sort(aValueUpperBound); mSorted = true; }
		void sort() { sort(hidden_list().requested_length()); }

		// ---- extensions used by the fused operators below ----
		gpu_list<4>& index_list() { return mIndexList; }
		void mark_sorted() { mSorted = true; }
		/// the other indexed_lists sharing the hidden data (they need the generic re-map when the hidden list is permuted)
		std::vector<indexed_list*> sharers() { std::vector<indexed_list*> v; for (auto* o : mHiddenData->mOwners) if (o != this) v.push_back(o); return v; }
		void apply_hidden_edit_public(gpu_list<4>& aEditList) { apply_hidden_edit(aEditList); }
		/// a list of ALL hidden entries (the scene's mParticles, pool.cpp:7) after copies were appended to the hidden list: 0 .. length-1
		void follow_hidden_growth()
		{
			mIndexList.write();
			mIndexList.set_length(mHiddenData->mData.length());
			shader_provider::write_sequence(mIndexList.buffer(), mIndexList.length(), 0u, 1u);
			mSorted = true;
		}

	private:
		class hidden_data : public list_interface<gpu_list<4>>
		{
		public:
			hidden_data() { mData.set_owner(this); }
			void apply_edit(gpu_list<4>& aEditList, owner_ptr aEditSource) override
			{
				for (auto* o : mOwners) if (static_cast<owner_ptr>(o) != aEditSource) o->apply_hidden_edit(aEditList);
				if (static_cast<owner_ptr>(&mData) != aEditSource) mData.apply_edit(aEditList, this);
			}
			DataList mData;
			std::list<indexed_list*> mOwners;
		};

		/// the hidden list is about to become hidden'[h] = hidden[edit[h]]: point every index at the new slot(s) of its
		/// entry, in ascending order (the state the reference reaches after its follow-up sort()), and let the owner follow
		void apply_hidden_edit(gpu_list<4>& aEditList)
		{
			if (empty()) return;
			auto newEditList = gpu_list<4>().request_length(mIndexList.requested_length());
			newEditList.set_length(0);
			auto old = mIndexList.rewrite_begin();
			shader_provider::check(apbf_apply_hidden_edit(shader_provider::context(), aEditList.buffer().template as<uint32_t>(), aEditList.length().template as<uint32_t>(), aEditList.capacity(),
			                                              old->mBuffer.template as<uint32_t>(), old->mLength.template as<uint32_t>(), static_cast<uint32_t>(old->mBuffer.bytes / 4),
			                                              static_cast<uint32_t>(mHiddenData->mData.requested_length()), mIndexList.buffer().template as<uint32_t>(),
			                                              newEditList.write().buffer().template as<uint32_t>(), mIndexList.length().template as<uint32_t>()));
			mSorted = true;
			if (mOwner == nullptr) return;
			newEditList.set_length(mIndexList.length());
			mOwner->apply_edit(newEditList, this);
		}

		gpu_list<4> mIndexList;
		std::shared_ptr<hidden_data> mHiddenData;
		owner_ptr mOwner = nullptr;
		bool mSorted = true;
	};

	// ---- source/list_definitions.h:9-20 ----
	enum class hidden_particles_enum { position, velocity, inverse_mass, radius, pos_backup, transferring };
	using hidden_particles = uninterleaved_list<hidden_particles_enum, gpu_list<16>, gpu_list<16>, gpu_list<4>, gpu_list<4>, gpu_list<16>, gpu_list<4>>;
	using particles = indexed_list<hidden_particles>;
	enum class fluid_enum { particle, target_radius, kernel_width, boundariness, boundary_distance };
	using fluid = uninterleaved_list<fluid_enum, particles, gpu_list<4>, gpu_list<4>, gpu_list<4>, gpu_list<4>>;
	using neighbors = gpu_list<8>;
	enum class hidden_transfers_enum { source, target, time_left };
	using hidden_transfers = uninterleaved_list<hidden_transfers_enum, particles, particles, gpu_list<4>>;
	using transfers = indexed_list<hidden_transfers>;

	/// runtime settings (source/settings.h static globals -> the context's apbf_settings + DIMENSIONS)
	class settings
	{
	public:
		static inline float splitDuration = 0.0f; // settings.cpp:23 (host-side: update_transfers.cpp:63)
		static void update_apbf_settings_buffer(const apbf_settings& aSettings, int aDimensions)
		{
			shader_provider::check(apbf_ctx_set_settings(shader_provider::context(), &aSettings));
			shader_provider::check(apbf_ctx_set_dimensions(shader_provider::context(), aDimensions));
		}
	};

	namespace detail
	{
		/// views of the reference's lists for one fused C-ABI call
		struct reorder_scope
		{
			std::vector<std::shared_ptr<gpu_list_data>> keep; // previous storages stay alive until the call has been enqueued
			template<size_t S> apbf_array rewrite(gpu_list<S>& l)
			{
				apbf_array a{ nullptr, nullptr };
				if (l.empty()) return a;
				auto old = l.rewrite_begin();
				a.data = old->mBuffer.ptr;
				a.reorder_out = l.buffer().ptr;
				keep.push_back(old);
				return a;
			}
		};
		template<size_t S> inline apbf_array in_place(gpu_list<S>& l)
		{
			apbf_array a{ nullptr, nullptr };
			if (!l.write().empty()) a.data = l.buffer().ptr;
			return a;
		}
		inline fluid* owning_fluid(particles* p) { return p ? dynamic_cast<fluid*>(p->owner()) : nullptr; }
		inline void fill_particles_in_place(particles& p, apbf_particles& o)
		{
			using hp = hidden_particles_enum;
			auto& h = p.hidden_list();
			o.index_list = in_place(p.index_list());
			o.length = p.length().as<uint32_t>();
			o.capacity = p.index_list().capacity();
			o.position = in_place(h.get<hp::position>());
			o.velocity = in_place(h.get<hp::velocity>());
			o.inverse_mass = in_place(h.get<hp::inverse_mass>());
			o.radius = in_place(h.get<hp::radius>());
			o.pos_backup = in_place(h.get<hp::pos_backup>());
			o.transferring = in_place(h.get<hp::transferring>());
			o.hidden_length = h.length().as<uint32_t>();
			o.hidden_capacity = h.get<hp::position>().capacity();
		}
		inline void fill_fluid_in_place(fluid& f, apbf_fluid& o)
		{
			std::memset(&o, 0, sizeof o);
			fill_particles_in_place(f.get<fluid_enum::particle>(), o.particle);
			o.target_radius = in_place(f.get<fluid_enum::target_radius>());
			o.kernel_width = in_place(f.get<fluid_enum::kernel_width>());
			o.boundariness = in_place(f.get<fluid_enum::boundariness>());
			o.boundary_distance = in_place(f.get<fluid_enum::boundary_distance>());
		}
		inline apbf_neighbors neighbors_view(neighbors& n)
		{
			n.make_unique();
			if (n.empty()) throw std::runtime_error("the neighbour list has no requested length");
			return apbf_neighbors{ n.buffer().as<uint32_t>(), n.length().as<uint32_t>(), n.capacity() };
		}

		/// shared by both searches: every list the search re-orders moves to fresh storage (copy-on-write, like the
		/// apply_edit chain of SURVEY 3.5), one fused call does the work, then the lengths follow.
		template<class Call>
		inline void run_search(particles* aParticles, const gpu_list<4>* aRange, neighbors* aNeighbors, Call&& aCall)
		{
			using hp = hidden_particles_enum;
			if (!aParticles || !aRange || !aNeighbors) throw std::runtime_error("neighborhood search: set_data() has not been called");
			fluid* fl = owning_fluid(aParticles);
			auto& hidden = aParticles->hidden_list();
			if (hidden.empty() || aParticles->empty()) { aNeighbors->set_length(0); return; }
			// Other indexed lists that share the hidden particles (the scene's mParticles, the transfers' source/target lists, ...)
			// are re-mapped by the generic path (indexed_list::apply_hidden_edit) with the permutation the search reports.
			auto sharers = aParticles->sharers();
			gpu_list<4> hiddenEdit;
			apbf_search_debug dbg;
			std::memset(&dbg, 0, sizeof dbg);
			if (!sharers.empty()) {
				hiddenEdit.request_length(hidden.get<hp::position>().capacity());
				hiddenEdit.set_length(hidden.length());
				dbg.sorted_index = hiddenEdit.write().buffer().as<uint32_t>();
			}
			reorder_scope scope;
			apbf_fluid f;
			std::memset(&f, 0, sizeof f);
			apbf_particles& p = f.particle;
			const auto oldIndexLength = aParticles->length();   // device words of the old storages stay valid while `scope` lives
			const auto oldHiddenLength = hidden.length();
			p.capacity = aParticles->index_list().capacity();
			p.hidden_capacity = hidden.get<hp::position>().capacity();
			p.index_list = scope.rewrite(aParticles->index_list());
			p.position = scope.rewrite(hidden.get<hp::position>());
			p.velocity = scope.rewrite(hidden.get<hp::velocity>());
			p.inverse_mass = scope.rewrite(hidden.get<hp::inverse_mass>());
			p.radius = scope.rewrite(hidden.get<hp::radius>());
			p.pos_backup = scope.rewrite(hidden.get<hp::pos_backup>());
			p.transferring = scope.rewrite(hidden.get<hp::transferring>());
			p.length = oldIndexLength.as<uint32_t>();
			p.hidden_length = oldHiddenLength.as<uint32_t>();
			apbf_array range{ nullptr, nullptr };
			bool rangeIsMember = false;
			if (fl) {
				auto take = [&](gpu_list<4>& l) { const bool isRange = (&l == aRange); auto a = scope.rewrite(l); if (isRange) { range = a; rangeIsMember = true; } return a; };
				f.target_radius = take(fl->get<fluid_enum::target_radius>());
				f.kernel_width = take(fl->get<fluid_enum::kernel_width>());
				f.boundariness = take(fl->get<fluid_enum::boundariness>());
				f.boundary_distance = take(fl->get<fluid_enum::boundary_distance>());
			}
			gpu_list<4> rangeScratch; // a range list outside the fluid is read by id and left alone, like in the reference
			if (!rangeIsMember) {
				rangeScratch = *aRange;
				range = scope.rewrite(rangeScratch);
			}
			apbf_neighbors nb = neighbors_view(*aNeighbors);
			aCall(&f, &range, &nb, sharers.empty() ? nullptr : &dbg);
			aParticles->mark_sorted();
			for (auto* other : sharers) other->apply_hidden_edit_public(hiddenEdit);
		}
	}

	class neighborhood_green
	{
	public:
		neighborhood_green& set_data(particles* aParticles, const gpu_list<sizeof(float)>* aRange, pbd::neighbors* aNeighbors) { mParticles = aParticles; mRange = aRange; mNeighbors = aNeighbors; return *this; }
		neighborhood_green& set_range_scale(float aScale) { mRangeScale = aScale; return *this; }
		neighborhood_green& set_position_range(const vec3& aMinPos, const vec3& aMaxPos, uint32_t aResolutionLog2) { mMinPos = aMinPos; mMaxPos = aMaxPos; mResolutionLog2 = aResolutionLog2; return *this; }
		void apply()
		{
			const float mn[3] = { mMinPos.x, mMinPos.y, mMinPos.z }, mx[3] = { mMaxPos.x, mMaxPos.y, mMaxPos.z };
			detail::run_search(mParticles, mRange, mNeighbors, [&](apbf_fluid* f, apbf_array* r, apbf_neighbors* n, apbf_search_debug* d) {
				shader_provider::check(apbf_neighborhood_green_apply(shader_provider::context(), f, r, n, mRangeScale, mn, mx, mResolutionLog2, d));
			});
		}
	private:
		float mRangeScale = 1.0f;
		particles* mParticles = nullptr;
		const gpu_list<sizeof(float)>* mRange = nullptr;
		pbd::neighbors* mNeighbors = nullptr;
		vec3 mMinPos, mMaxPos;
		uint32_t mResolutionLog2 = 4u;
	};

	class neighborhood_binary_search
	{
	public:
		neighborhood_binary_search& set_data(particles* aParticles, const gpu_list<sizeof(float)>* aRange, pbd::neighbors* aNeighbors) { mParticles = aParticles; mRange = aRange; mNeighbors = aNeighbors; return *this; }
		neighborhood_binary_search& set_range_scale(float aScale) { mRangeScale = aScale; return *this; }
		void apply()
		{
			detail::run_search(mParticles, mRange, mNeighbors, [&](apbf_fluid* f, apbf_array* r, apbf_neighbors* n, apbf_search_debug* d) {
				shader_provider::check(apbf_neighborhood_binary_search_apply(shader_provider::context(), f, r, n, mRangeScale, d));
			});
		}
	private:
		float mRangeScale = 1.0f;
		particles* mParticles = nullptr;
		const gpu_list<sizeof(float)>* mRange = nullptr;
		pbd::neighbors* mNeighbors = nullptr;
	};

	class incompressibility
	{
	public:
		incompressibility& set_data(fluid* aFluid, neighbors* aNeighbors) { mFluid = aFluid; mNeighbors = aNeighbors; return *this; }
		void apply()
		{
			if (!mFluid || !mNeighbors) throw std::runtime_error("incompressibility: set_data() has not been called");
			if (mFluid->empty()) return;
			apbf_fluid f;
			detail::fill_fluid_in_place(*mFluid, f);
			apbf_neighbors nb = detail::neighbors_view(*mNeighbors);
			shader_provider::check(apbf_incompressibility_apply(shader_provider::context(), &f, &nb, nullptr, nullptr));
		}
	private:
		fluid* mFluid = nullptr;
		neighbors* mNeighbors = nullptr;
	};

	class spread_kernel_width
	{
	public:
		spread_kernel_width& set_data(fluid* aFluid, neighbors* aNeighbors) { mFluid = aFluid; mNeighbors = aNeighbors; return *this; }
		void apply()
		{
			if (!mFluid || !mNeighbors) throw std::runtime_error("spread_kernel_width: set_data() has not been called");
			if (mFluid->empty()) return;
			apbf_fluid f;
			detail::fill_fluid_in_place(*mFluid, f);
			apbf_neighbors nb = detail::neighbors_view(*mNeighbors);
			shader_provider::check(apbf_spread_kernel_width_apply(shader_provider::context(), &f, &nb, nullptr));
		}
	private:
		fluid* mFluid = nullptr;
		neighbors* mNeighbors = nullptr;
	};

	namespace detail
	{
		/// the reference's transfer lists (list_definitions.h:16-18) as the C-ABI's view; source and target are index lists into the
		/// hidden particle data (pool.cpp:18-19)
		inline void check_transfers(transfers& t)
		{
			using ht = hidden_transfers_enum;
			auto& h = t.hidden_list();
			if (h.get<ht::source>().index_list().requested_length() == 0 || h.get<ht::target>().index_list().requested_length() == 0 || h.get<ht::time_left>().requested_length() == 0)
				throw std::runtime_error("transfers: the hidden list has no requested length (construct with transfers(MAX_TRANSFERS))");
		}
	}

	/// pbd::update_transfers (source/update_transfers.h, update_transfers.cpp:14-70).  Without a transfers list (or with
	/// mMerge and mSplit off) only the first part runs (boundary-distance flood fill, nearest neighbour, target radius); with one,
	/// merges are entered and splits are started as the context's settings say (conflicts resolved in ascending id order).
	class update_transfers
	{
	public:
		update_transfers& set_data(fluid* aFluid, neighbors* aNeighbors, transfers* aTransfers = nullptr) { mFluid = aFluid; mNeighbors = aNeighbors; mTransfers = aTransfers; return *this; }
		void apply()
		{
			if (!mFluid || !mNeighbors) throw std::runtime_error("update_transfers: set_data() has not been called");
			if (mFluid->empty()) return;
			apbf_fluid f;
			detail::fill_fluid_in_place(*mFluid, f);
			apbf_neighbors nb = detail::neighbors_view(*mNeighbors);
			if (!mTransfers) {
				shader_provider::check(apbf_update_transfers_apply(shader_provider::context(), &f, &nb, nullptr));
				return;
			}
			using ht = hidden_transfers_enum;
			using hp = hidden_particles_enum;
			detail::check_transfers(*mTransfers);
			auto& rows = mTransfers->hidden_list();
			auto& source = rows.get<ht::source>();
			auto& target = rows.get<ht::target>();
			auto& timeLeft = rows.get<ht::time_left>();
			apbf_transfers t;
			std::memset(&t, 0, sizeof t);
			t.source = detail::in_place(source.index_list());
			t.target = detail::in_place(target.index_list());
			t.time_left = detail::in_place(timeLeft);
			t.length = source.length().as<uint32_t>();
			t.capacity = std::min(source.index_list().capacity(), std::min(target.index_list().capacity(), timeLeft.capacity()));
			shader_provider::check(apbf_update_transfers_split_merge_apply(shader_provider::context(), &f, &nb, &t, settings::splitDuration, nullptr));
			// every member list carries its own length word: the new lengths reach the siblings like in update_transfers.cpp:54
			auto& particleList = mFluid->get<fluid_enum::particle>();
			auto& hidden = particleList.hidden_list();
			hidden.set_length(hidden.get<hp::position>().length());
			mFluid->set_length(particleList.length());
			rows.set_length(source.length());
			for (auto* other : particleList.sharers())   // the scene's own list of all particles gains the copies as well
				if (other != &source && other != &target && other->owner() == nullptr && other->index_list().capacity() > 0) other->follow_hidden_growth();
		}
	private:
		fluid* mFluid = nullptr;
		neighbors* mNeighbors = nullptr;
		transfers* mTransfers = nullptr;
	};

	/// pbd::particle_transfer (source/particle_transfer.h, particle_transfer.cpp:10-28): one time step of every transfer under way;
	/// finished splits leave the transfer list, the sources of finished merges leave the particle lists (surviving entries keep
	/// their order)
	class particle_transfer
	{
	public:
		particle_transfer& set_data(fluid* aFluid, transfers* aTransfers) { mFluid = aFluid; mTransfers = aTransfers; return *this; }
		void apply(float aDeltaTime)
		{
			if (!mFluid || !mTransfers) throw std::runtime_error("particle_transfer: set_data() has not been called");
			using ht = hidden_transfers_enum;
			using hp = hidden_particles_enum;
			auto& particleList = mFluid->get<fluid_enum::particle>();
			auto& hidden = particleList.hidden_list();
			if (hidden.empty() || particleList.empty()) return;
			detail::check_transfers(*mTransfers);
			auto& rows = mTransfers->hidden_list();
			auto& source = rows.get<ht::source>();
			auto& target = rows.get<ht::target>();
			auto& timeLeft = rows.get<ht::time_left>();
			source.index_list().write(); target.index_list().write(); timeLeft.write();
			// other lists sharing the hidden particles follow through the generic path with the edit the call reports
			std::vector<particles*> others;
			for (auto* o : particleList.sharers()) if (o != &source && o != &target) others.push_back(o);
			gpu_list<4> hiddenEdit;
			uint32_t* editOut = nullptr;
			if (!others.empty()) {
				hiddenEdit.request_length(hidden.get<hp::position>().capacity());
				hiddenEdit.set_length(0);
				editOut = hiddenEdit.write().buffer().as<uint32_t>();
			}
			detail::reorder_scope scope;
			apbf_fluid f;
			std::memset(&f, 0, sizeof f);
			apbf_particles& p = f.particle;
			p.capacity = particleList.index_list().capacity();
			p.hidden_capacity = hidden.get<hp::position>().capacity();
			p.index_list = scope.rewrite(particleList.index_list());
			p.position = scope.rewrite(hidden.get<hp::position>());
			p.velocity = scope.rewrite(hidden.get<hp::velocity>());
			p.inverse_mass = scope.rewrite(hidden.get<hp::inverse_mass>());
			p.radius = scope.rewrite(hidden.get<hp::radius>());
			p.pos_backup = scope.rewrite(hidden.get<hp::pos_backup>());
			p.transferring = scope.rewrite(hidden.get<hp::transferring>());
			f.target_radius = scope.rewrite(mFluid->get<fluid_enum::target_radius>());
			f.kernel_width = scope.rewrite(mFluid->get<fluid_enum::kernel_width>());
			f.boundariness = scope.rewrite(mFluid->get<fluid_enum::boundariness>());
			f.boundary_distance = scope.rewrite(mFluid->get<fluid_enum::boundary_distance>());
			// the new storages start with a copy of the old length words; the call reads them and leaves the new lengths there
			p.length = particleList.length().as<uint32_t>();
			p.hidden_length = hidden.get<hp::position>().length().as<uint32_t>();
			apbf_transfers t;
			std::memset(&t, 0, sizeof t);
			t.source = scope.rewrite(source.index_list());
			t.target = scope.rewrite(target.index_list());
			t.time_left = scope.rewrite(timeLeft);
			t.length = source.length().as<uint32_t>();
			t.capacity = std::min(source.index_list().capacity(), std::min(target.index_list().capacity(), timeLeft.capacity()));
			shader_provider::check(apbf_particle_transfer_apply(shader_provider::context(), &f, &t, aDeltaTime, editOut));
			hidden.set_length(hidden.get<hp::position>().length());
			mFluid->set_length(particleList.length());
			rows.set_length(source.length());
			if (!others.empty()) {
				hiddenEdit.set_length(hidden.get<hp::position>().length());
				for (auto* o : others) o->apply_hidden_edit_public(hiddenEdit);
			}
		}
	private:
		fluid* mFluid = nullptr;
		transfers* mTransfers = nullptr;
	};

	/// pbd::save_particle_info::apply() (source/save_particle_info.cpp:21-130): the reference's on-disk particle dump --
	/// six ';'-separated text files and data.csv (sorted by the distance to (0, 10, -60)) in PARTICLE_INFO_FOLDER_NAME
	/// ("particle_data", cpu_gpu_shared_config.h:21), floats in the default ostream format.  Host-side only; lets a run of
	/// this library be diffed against a run of the reference.  save_as_svg is rendering and not provided.
	class save_particle_info
	{
	public:
		save_particle_info& set_data(fluid* aFluid, neighbors* aNeighbors, transfers* aTransfers = nullptr) { mFluid = aFluid; mNeighbors = aNeighbors; (void)aTransfers; return *this; }
		save_particle_info& set_folder(const std::string& aFolder) { mFolder = aFolder; return *this; } // default: the reference's
		void apply()
		{
			if (!mFluid || !mNeighbors) throw std::runtime_error("save_particle_info: set_data() has not been called");
			using hp = hidden_particles_enum;
			auto& parts = mFluid->get<fluid_enum::particle>();
			auto indices = parts.index_read();
			auto positions = parts.hidden_list().template get<hp::position>().template read<int32_t>();   // ivec4 per particle
			auto radii = parts.hidden_list().template get<hp::radius>().template read<float>();
			auto invMasses = parts.hidden_list().template get<hp::inverse_mass>().template read<float>();
			auto kernelWidth = mFluid->get<fluid_enum::kernel_width>().template read<float>();
			auto targetRadius = mFluid->get<fluid_enum::target_radius>().template read<float>();
			auto boundaryDist = mFluid->get<fluid_enum::boundary_distance>().template read<uint32_t>();
			auto nbrPairs = mNeighbors->template read<uint32_t>();                                         // uvec2 per pair
			const size_t n = indices.size();
			std::vector<unsigned int> nbrCount(n, 0u), sortedIdx(n);
			std::vector<float> radius(n), inverseMass(n), centerDist(n), bdrDist(n), px(n), py(n), pz(n);
			for (size_t e = 0; e + 1 < nbrPairs.size(); e += 2) nbrCount[nbrPairs[e]]++;
			for (size_t i = 0; i < n; i++) {
				const uint32_t id = indices[i];
				px[i] = static_cast<float>(positions[4 * id]) / APBF_POS_RESOLUTION;
				py[i] = static_cast<float>(positions[4 * id + 1]) / APBF_POS_RESOLUTION;
				pz[i] = static_cast<float>(positions[4 * id + 2]) / APBF_POS_RESOLUTION;
				bdrDist[i] = boundaryDist[i] / APBF_POS_RESOLUTION;
				const float dx = px[i] - 0.0f, dy = py[i] - 10.0f, dz = pz[i] - (-60.0f);              // centerPos, :49
				centerDist[i] = std::sqrt(dx * dx + dy * dy + dz * dz);
				radius[i] = radii[id];
				inverseMass[i] = invMasses[id];
				sortedIdx[i] = static_cast<unsigned int>(i);
			}
			std::filesystem::create_directories(mFolder);
			auto dump = [&](const char* name, const auto& v) { std::ofstream f(mFolder + "/" + name); for (auto& x : v) f << x << ";"; };
			dump("centerDist.txt", centerDist);
			dump("radius.txt", radius);
			dump("neighborCount.txt", nbrCount);
			dump("kernelWidth.txt", kernelWidth);
			dump("targetRadius.txt", targetRadius);
			dump("boundaryDistance.txt", bdrDist);
			std::sort(sortedIdx.begin(), sortedIdx.end(), [&](unsigned int a, unsigned int b) { return centerDist[a] < centerDist[b]; });
			std::ofstream f(mFolder + "/data.csv");
			f << "center distance,boundary distance,kernel width,neighbor count,radius,target radius,inverse mass,x,y,z" << std::endl;
			for (size_t i = 0; i < n; i++) {
				const unsigned int k = sortedIdx[i];
				f << centerDist[k] << "," << bdrDist[k] << "," << kernelWidth[k] << "," << nbrCount[k] << "," << radius[k] << ","
				  << targetRadius[k] << "," << inverseMass[k] << "," << px[k] << "," << py[k] << "," << pz[k] << std::endl;
			}
		}
	private:
		fluid* mFluid = nullptr;
		neighbors* mNeighbors = nullptr;
		std::string mFolder = "particle_data";
	};

	class box_collision
	{
	public:
		box_collision& set_data(particles* aParticles, pbd::gpu_list<16>* aBoxMin, pbd::gpu_list<16>* aBoxMax) { mParticles = aParticles; mBoxMin = aBoxMin; mBoxMax = aBoxMax; return *this; }
		/// the number of boxes is host-side knowledge of the caller (user_controlled_boxes owns the lists); default: the capacity of aBoxMin
		box_collision& set_box_count(uint32_t aCount) { mBoxCount = static_cast<int64_t>(aCount); return *this; }
		void apply()
		{
			if (!mParticles || !mBoxMin || !mBoxMax) throw std::runtime_error("box_collision: set_data() has not been called");
			if (mParticles->empty()) return;
			apbf_particles p;
			std::memset(&p, 0, sizeof p);
			detail::fill_particles_in_place(*mParticles, p);
			const uint32_t n = mBoxMin->empty() ? 0u : (mBoxCount >= 0 ? static_cast<uint32_t>(mBoxCount) : mBoxMin->capacity());
			shader_provider::check(apbf_box_collision_apply(shader_provider::context(), &p, n ? mBoxMin->buffer().as<float>() : nullptr, n ? mBoxMax->buffer().as<float>() : nullptr, n));
		}
	private:
		particles* mParticles = nullptr;
		pbd::gpu_list<16>* mBoxMin = nullptr;
		pbd::gpu_list<16>* mBoxMax = nullptr;
		int64_t mBoxCount = -1;
	};

	class velocity_handling
	{
	public:
		velocity_handling& set_data(particles* aParticles) { mParticles = aParticles; return *this; }
		velocity_handling& set_acceleration(const vec3& aAcceleration = vec3(0)) { mAcceleration = aAcceleration; return *this; }
		void apply(float aDeltaTime)
		{
			if (!mParticles) throw std::runtime_error("velocity_handling: set_data() has not been called");
			if (mParticles->empty()) return;
			apbf_particles p;
			std::memset(&p, 0, sizeof p);
			detail::fill_particles_in_place(*mParticles, p);
			const float a[3] = { mAcceleration.x, mAcceleration.y, mAcceleration.z };
			shader_provider::check(apbf_velocity_handling_apply(shader_provider::context(), &p, aDeltaTime, mLastDeltaTime, a));
			if (aDeltaTime != 0.0f) mLastDeltaTime = aDeltaTime;
		}
	private:
		particles* mParticles = nullptr;
		vec3 mAcceleration;
		float mLastDeltaTime = 1.0f;
	};
}
